#!/usr/bin/env python
"""Kernel lab for mnv_step: graph-replayed period per launch for several batch sizes and option settings, plus a quick
teacher-forced parity check against the oracle.  Development tool (GPU box only); bench.py is the contract.

    python scripts/step_lab.py --envs 64,4736,65536,262144 --opts "pdl=0;pdl=1"
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from distributional_rl_navigation_b200 import _lib, env_ops
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv


def set_opts(spec):
    L = _lib.load()
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        rc = L.mnv_set_option(k.encode(), int(v))
        assert rc == 0, (k, v, rc)


def period_us(E, steps, n_c=4, n_o=8, n_b=11, dev="cuda:0", n_sub=None, drift=False, n_streams=1):
    per_env = 490 if n_b == 11 else 1490
    nb = max(2, min(8, int(300e6 // (E * per_env)) + 1))
    batches = []
    for b in range(nb):
        env = VecMarineNavEnv(E, seed=b * E, device=dev, num_cores=n_c, num_obs=n_o, min_start_goal_dis=30.0, num_beams=n_b)
        env.reset()
        env.rng_key = None
        batches.append(env)
    g = torch.Generator(device=dev); g.manual_seed(1)
    actions = torch.randint(0, 9, (64, E), generator=g, device=dev, dtype=torch.int32)
    params = batches[0].params()
    if n_sub is not None:
        params.n_substeps = n_sub

    def one(i):
        env_ops.step(batches[i % nb].buf, params, action=actions[i % 64])
    for i in range(8):
        one(i)
    torch.cuda.synchronize()
    chunk = 64
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    extra = [torch.cuda.Stream() for _ in range(n_streams - 1)]
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            # n_streams > 1: launch i goes to stream i % n_streams (the batches are independent), so that the CTAs of the
            # next launch fill the SM slots the draining launch frees instead of waiting at the grid dependency
            for s_ in extra:
                s_.wait_stream(side)
            lanes = [side] + extra
            for i in range(chunk):
                with torch.cuda.stream(lanes[i % n_streams]):
                    one(i)
            for s_ in extra:
                side.wait_stream(s_)
    torch.cuda.current_stream().wait_stream(side)
    reps = max(1, steps // chunk)
    if drift:
        # period of single replays as the (never reset) robots drift away from the stationary rollout distribution
        out = []
        done_steps = 8
        for r in range(2049):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); graph.replay(); e1.record()
            torch.cuda.synchronize()
            done_steps += chunk
            if r in (0, 1, 2, 3, 7, 15, 31, 63, 127, 255, 511, 1023, 2047):
                frac_done = float(sum(int((b.buf["done"] != 0).sum()) for b in batches)) / (nb * E)
                out.append((done_steps // nb, round(e0.elapsed_time(e1) * 1e3 / chunk, 2), round(frac_done, 3)))
        print("   drift (steps per batch since reset, us per step, fraction of envs flagged done):", out, flush=True)
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (reps * chunk))
    del batches, graph
    torch.cuda.empty_cache()
    return best, nb


def parity(E=8192, n_steps=4, dev="cuda:0"):
    from oracle import marinenav_oracle as mo
    op = mo.default_params(11)
    w = mo.reset_batch(np.arange(E, dtype=np.uint32) + 77, 4, 8, 30.0, 4, 8, op, n_threads=8)
    p = _lib.default_params(11)
    buf = env_ops.alloc_env_buffers(E, 4, 8, 11, dev)
    for k in ("state", "velocity", "goal", "cores", "obstacles"):
        buf[k].copy_(torch.from_numpy(w[k]))
    rng = np.random.RandomState(0)
    ep = np.zeros(E, np.int32)
    worst = dict(obs=0.0, rew=0.0, state=0.0, flags=0)
    for t in range(n_steps):
        action = rng.randint(0, 9, size=E).astype(np.int32)
        buf["action"].copy_(torch.from_numpy(action))
        env_ops.step(buf, p)
        obs, reward, done, info = mo.step_batch(w["state"], w["velocity"], w["goal"], w["cores"], w["obstacles"], action, ep, op, 8)
        torch.cuda.synchronize()
        g = buf["obs"].cpu().numpy()
        worst["obs"] = max(worst["obs"], float((np.abs(g - obs) / np.maximum(1.0, np.abs(obs))).max()))
        worst["rew"] = max(worst["rew"], float(np.abs(buf["reward"].cpu().numpy() - reward).max()))
        worst["state"] = max(worst["state"], float(np.abs(buf["state"].cpu().numpy() - w["state"]).max()))
        worst["flags"] += int((buf["done"].cpu().numpy() != done).sum() + (buf["info"].cpu().numpy() != info).sum())
    return worst


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", default="65536")
    ap.add_argument("--opts", default="", help="';'-separated option sets, each 'k=v,k=v' (mnv_set_option)")
    ap.add_argument("--steps", type=int, default=1024)
    ap.add_argument("--dense", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--drift", action="store_true")
    ap.add_argument("--streams", default="1", help="comma list: launches distributed round-robin over this many streams")
    ap.add_argument("--shape", default="", help="';'-separated 'n_sub,n_beams' overrides (timing only)")
    a = ap.parse_args()
    for spec in a.opts.split(";"):
        set_opts(spec)
        line = f"[{spec or 'default'}]"
        if not a.no_parity:
            line += f" parity={parity()}"
        print(line, flush=True)
        for shp in filter(None, a.shape.split(";")):
            ns, nbm = [int(v) for v in shp.split(",")]
            for E in [int(x) for x in a.envs.split(",")]:
                us, nb = period_us(E, a.steps, 4, 32 if a.dense else 8, nbm, n_sub=ns)
                print(f"   shape n_sub={ns} n_beams={nbm} E={E:7d} period={us:8.2f} us", flush=True)
        for E, ns_ in [(int(x), int(y)) for x in a.envs.split(",") for y in a.streams.split(",")]:
            us, nb = period_us(E, a.steps, *((4, 32, 64) if a.dense else (4, 8, 11)), drift=a.drift, n_streams=ns_)
            per = 1490 if a.dense else 490
            print(f"   E={E:7d} streams={ns_} batches={nb} period={us:8.2f} us  {E / us / 1e3:7.3f} G steps/s  {E * per / us / 1e3 / 6532.5 * 100:5.1f}% HBM", flush=True)


if __name__ == "__main__":
    main()
