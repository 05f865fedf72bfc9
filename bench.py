#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for the marinenav hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one fused mnv_step launch over one batch of 65 536 environments per GPU (BASELINE configs[1]:
8 obstacles / 4 vortex cores / 11 beams, synthetic randomly seeded maps).  Consecutive steps rotate over 8 independent
env batches (8 x ~32 MB > 126 MB L2) so every step streams its state from HBM.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "batched env steps/sec and IQN updates/sec at 1/2/4/8 B200 vs CPU ref"
UNIT = "env_steps/s"
N_CORES, N_OBS, N_BEAMS = 4, 8, 11
ENVS_PER_GPU = 65536
N_BATCHES = 8
NCU_DRAM_BYTES_PER_LAUNCH = 22.589e6     # measured, profiles/r2_step_kernel_v8_ncu_full.csv (algorithmic read volume: 22.5 MB)
NCU_WARP_INST_PER_LAUNCH = 5.43e6        # smsp__inst_executed.sum of the same capture
NCU_TRAFFIC_SOURCE = ("dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture "
                      "(profiles/r2_step_kernel_v8_ncu_full.csv); the 9.6 MB of writes were still in L2 when the profiled launch ended")


def algorithmic_bytes_per_env_step(n_c=N_CORES, n_o=N_OBS, n_b=N_BEAMS, s=8):
    """read = s*(4 state + 2 goal + 3 n_o + 3 n_c) + 4 (timestep) + 4 (action);
    write = s*4 + 4 + 4*(4 + 2 n_b) + 4 (reward) + 2 (done, info).  fp64 tables -> s = 8 -> 490 B at C2."""
    return s * (4 + 2 + 3 * n_o + 3 * n_c) + 8 + (s * 4 + 4 + 4 * (4 + 2 * n_b) + 4 + 2)


def bench_config(E, world):
    """The workload description; BOTH arms (--impl b200 / reference) print exactly this dict as `config`."""
    abytes = algorithmic_bytes_per_env_step()
    return {"workload": "marinenav fused env step (MarineNavEnv.step), 65536 envs/GPU, 8 obstacles / 4 vortex cores / "
                        "11 beams, random actions (BASELINE configs[1])",
            "envs_per_gpu": E,
            "l2": f"rotating {N_BATCHES} env batches ({N_BATCHES * E * abytes / 1e6:.0f} MB > 126 MB L2), one batch per step; the steps of "
                  "independent batches are launched on two streams (a batch stays on one stream)",
            "auto_reset": "in e2e only; every batch is reset and mixed with auto-reset steps right before the timed region",
            "parallelism": f"env-sharded x{world}, no collective in step"}


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self._stop, self._t = gpu_index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True); self._t.start(); return self

    def __exit__(self, *a):
        self._stop.set(); self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][2]) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


_JSON_OUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line: keep a private handle to the real stdout and point fd 1 at stderr, so that
    library chatter (e.g. NCCL's version banner, printed with printf from C) cannot land in front of it."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle port; the only place besides tests/ and smoke() that executes oracle/)
# ---------------------------------------------------------------------------------------------------------------
def cpu_port_steps_per_s(n_envs, n_steps, n_threads, seed0=0):
    import numpy as np
    from oracle import marinenav_oracle as mo
    op = mo.default_params(N_BEAMS)
    w = mo.reset_batch(np.arange(n_envs, dtype=np.uint32) + seed0, N_CORES, N_OBS, 30.0, N_CORES, N_OBS, op, n_threads=n_threads)
    rng = np.random.RandomState(1)
    ep = np.zeros(n_envs, np.int32)
    acts = [rng.randint(0, 9, size=n_envs).astype(np.int32) for _ in range(n_steps)]
    t0 = time.perf_counter()
    for a in acts:
        mo.step_batch(w["state"], w["velocity"], w["goal"], w["cores"], w["obstacles"], a, ep, op, n_threads)
    dt = time.perf_counter() - t0
    return n_envs * n_steps / dt, dt


def cpu_baseline_leg(E, cores):
    """`cpu_baseline` of the B200 arm (rank 0, N = 1): the reference's own Python step() on all host cores (one process per
    core, staged in oracle/_ref) when available, with the C port beside it; bounded sample (~10-20 s of CPU work)."""
    from oracle import ref_timing
    port_v, port_dt = cpu_port_steps_per_s(E, 40, cores)
    port1_v, _ = cpu_port_steps_per_s(E, 20, 1)
    if not ref_timing.available():
        return {"value": port_v, "unit": UNIT, "cores": cores, "kind": "port", "value_1_thread": port1_v,
                "sample": f"{E} envs x 40 steps of the same workload, oracle/marinenav_oracle.c ({port_dt:.1f} s on {cores} threads); "
                          "oracle/_ref (the staged Python reference) is absent"}
    per = 1200
    ref_v, ref_dt = ref_timing.env_steps_per_s(cores, per, n_c=N_CORES, n_o=N_OBS, n_b=N_BEAMS)
    ref1_v, _ = ref_timing.env_steps_per_s(1, per, n_c=N_CORES, n_o=N_OBS, n_b=N_BEAMS)
    return {"value": ref_v, "unit": UNIT, "cores": cores, "kind": "reference", "value_1_core": ref1_v,
            "sample": f"{cores} processes x {per} MarineNavEnv.step calls (unmodified marinenav_env.py:199 staged in oracle/_ref, same map "
                      f"rules / random actions, reset on done; {ref_dt:.1f} s), and 1 process alone",
            "port": {"value": port_v, "value_1_thread": port1_v, "cores": cores,
                     "sample": f"{E} envs x 40 steps, oracle/marinenav_oracle.c (C restatement, {port_dt:.1f} s on {cores} threads)"}}


def params_pdl(p):
    return bool(getattr(p, "pdl_prefetch", 0))


def run_reference(args):
    """--impl reference: the reference's OWN CPU implementation of the path on the box's host cores.  When oracle/_ref
    (the staged, unmodified Python files: oracle/stage_ref.py) is present, P = cpu_count worker processes each step one
    MarineNavEnv (marinenav_env.py:199) -- kind "reference"; otherwise the C restatement on all host threads -- kind "port".
    Each timed step is a bounded sample of the 65 536-env workload (the Python reference needs ~10 s for one full step)."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import ref_timing
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    port_v, _ = cpu_port_steps_per_s(ENVS_PER_GPU, 5, cores)                # the C port beside it, same envs per step as the GPU arm
    if ref_timing.available():
        per = 96                                                             # env steps per worker per timed step (~0.3 s)
        durs = ref_timing.env_segments(cores, W + K, per, N_CORES, N_OBS, N_BEAMS)
        seg = [max(d[i] for d in durs) for i in range(W, W + K)]             # a step ends when the slowest worker is done
        total = sum(seg)
        value = cores * per * K / total
        kind, sample = "reference", (f"{cores} worker processes x {per} MarineNavEnv.step calls per timed step (unmodified "
                                     f"marinenav_env.py / robot.py staged in oracle/_ref, reset on done); C port of the same "
                                     f"path on {cores} threads at 65536 envs per step: {port_v:.3e} env-steps/s")
    else:
        seg = []
        for i in range(W + K):
            v, dt = cpu_port_steps_per_s(ENVS_PER_GPU, 1, cores, seed0=i)
            if i >= W:
                seg.append(dt)
        total = sum(seg)
        value = ENVS_PER_GPU * K / total
        kind, sample = "port", f"{ENVS_PER_GPU} envs x 1 step per timed step, oracle/marinenav_oracle.c on {cores} threads (oracle/_ref absent)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 * total / K, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(ENVS_PER_GPU, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "port_value": port_v},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from distributional_rl_navigation_b200 import _lib, env_ops
    from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv

    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    E, K, W = args.envs, args.steps, args.warmup
    W = max(W, 3)                                   # timing rules: at least 3 warm-up steps

    # ---- synthetic workload: N_BATCHES independent env batches, maps from the device reset (seed = global env index)
    batches = []
    for b in range(N_BATCHES):
        env = VecMarineNavEnv(E, seed=(rank * N_BATCHES + b) * E, device=dev, num_cores=N_CORES, num_obs=N_OBS,
                              min_start_goal_dis=30.0, num_beams=N_BEAMS)
        env.reset()
        batches.append(env)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    actions = torch.randint(0, 9, (W + K, E), generator=g, device=dev, dtype=torch.int32)
    params = batches[0].params()
    stream = torch.cuda.current_stream()

    def one_step(i):
        env_ops.step(batches[i % N_BATCHES].buf, params, action=actions[i % actions.shape[0]])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        one_step(i)
    barrier()
    # The step kernel lasts ~10 us, less than a Python-side launch, so the timed region consists of CUDA-graph replays ONLY:
    # one graph holds REPS x K consecutive mnv_step launches (through the C-ABI on the capturing stream; REPS chosen so that
    # a graph is >= 256 launches and a whole number of batch rotations), and it is replayed N_REPLAY times back to back with
    # an event after every replay, so that the region lasts >= 25 ms (30 ms at the pace of a first replay).  ms_per_step = median replay / (REPS x K).
    reps = max(1, -(-256 // K))
    while (reps * K) % N_BATCHES != 0 and reps < 4096:
        reps += 1
    n_launch = reps * K
    # The env batches are independent, so the launches alternate between N_STREAMS capture streams (launch i -> stream
    # i % N_STREAMS, batch i % N_BATCHES: a batch always stays on one stream, so its own steps remain ordered): every stream
    # is a chain of programmatic dependent launches, and the CTAs of the next launch fill the SM slots a draining launch
    # frees instead of idling at the grid dependency of a one-wave kernel.  `timing.single_stream` reports the same graph
    # captured on ONE stream beside it.
    n_streams = max(1, args.streams)
    assert N_BATCHES % n_streams == 0, "--streams must divide the number of rotating batches"
    side = torch.cuda.Stream(device=dev)
    lanes = [side] + [torch.cuda.Stream(device=dev) for _ in range(n_streams - 1)]

    def capture(n_lanes):
        gr = torch.cuda.CUDAGraph()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(gr, stream=side):
                for s_ in lanes[1:n_lanes]:
                    s_.wait_stream(side)
                for i in range(n_launch):
                    with torch.cuda.stream(lanes[i % n_lanes]):
                        one_step(W + i)
                for s_ in lanes[1:n_lanes]:
                    side.wait_stream(s_)
        torch.cuda.current_stream().wait_stream(side)
        return gr

    graph = capture(n_streams)
    graph_1s = capture(1) if n_streams > 1 else None
    stream = torch.cuda.current_stream()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graph.replay(); torch.cuda.synchronize()
    c0.record(stream); graph.replay(); c1.record(stream); torch.cuda.synchronize()
    n_replay = int(min(400, max(5, -(-30.0 // c0.elapsed_time(c1)))))

    def timed_region(evs=None):
        for r in range(n_replay):
            graph.replay()
            if evs is not None:
                evs[r + 1].record(stream)

    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_replay + 1)]
    with ClockSampler(local_rank) as clk:
        # ~1 s of the same kernels right before the timed region so that nvidia-smi sees the clocks under THIS load
        # (the timed region itself lasts tens of milliseconds); clocks are sampled across both.
        t_pre = time.perf_counter()
        while time.perf_counter() - t_pre < args.preload:
            timed_region()
            torch.cuda.synchronize()
        # The untimed preload above ran tens of thousands of steps without resets, so the robots have long left their
        # maps.  Put every batch back into the stationary rollout distribution (fresh maps, then MIX auto-reset steps of the
        # random policy: episodes in every phase, finished ones re-drawn) right before the timed region; the timed region
        # itself then steps each batch n_launch x n_replay / N_BATCHES times without resets.
        for bi, env in enumerate(batches):
            env.reset()
            for i in range(args.mix):
                env.step(actions[(bi + i) % actions.shape[0]], auto_reset=True)
        for i in range(N_BATCHES):                   # caches / instruction memory warm again after the reset kernels
            one_step(W + i)
        barrier()
        evs[0].record(stream)
        timed_region(evs)
        barrier()
        per_replay = sorted(evs[r].elapsed_time(evs[r + 1]) for r in range(n_replay))
        ms_region = evs[0].elapsed_time(evs[n_replay])
        ms_1s = None
        if graph_1s is not None:                     # the single-stream period of the same launches (reported, not the value)
            graph_1s.replay(); torch.cuda.synchronize()
            e1s = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            e1s[0].record(stream)
            for r in range(5):
                graph_1s.replay(); e1s[r + 1].record(stream)
            barrier()
            ms_1s = sorted(e1s[r].elapsed_time(e1s[r + 1]) for r in range(5))[2] / n_launch
    ms_median = per_replay[len(per_replay) // 2]
    t = torch.tensor([ms_median / n_launch, ms_region / (n_launch * n_replay)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step, ms_per_step_mean = float(t[0].item()), float(t[1].item())
    value = world * E / (ms_per_step * 1e-3)

    # ---- device-resident rollout step: fused step + masked reset + masked re-observe of the finished environments
    #      (VecMarineNavEnv.step(auto_reset=True)), one batch, CUDA graph of 16 steps ----
    env0 = batches[0]
    env0.reset()
    for i in range(args.mix):
        env0.step(actions[i % actions.shape[0]], auto_reset=True)
    g_ar = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g_ar, stream=side):
            for i in range(16):
                env0.step(actions[i % actions.shape[0]], auto_reset=True)
    torch.cuda.current_stream().wait_stream(side)
    g_ar.replay(); torch.cuda.synchronize()
    n_ar = 12
    ar = [torch.cuda.Event(enable_timing=True) for _ in range(n_ar + 1)]
    ar[0].record(stream)
    for r in range(n_ar):
        g_ar.replay(); ar[r + 1].record(stream)
    barrier()
    ar_ms = sorted(ar[r].elapsed_time(ar[r + 1]) / 16 for r in range(n_ar))[n_ar // 2]
    ta = torch.tensor([ar_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
    ar_ms = float(ta.item())
    finished_frac = float(env0.buf["done"].float().mean().item())

    # ---- e2e: public API with HOST buffers (H2D actions, D2H obs/reward/done/info, auto-reset) every step: K steps per
    #      repetition, repeated until >= ~0.1 s, median repetition ----
    host_actions = np.random.RandomState(7 + rank).randint(0, 9, size=(W + K, E)).astype(np.int32)
    for i in range(max(W, 100)):                        # host_transport="auto" measures the transports in calls 64 .. 96
        env0.step_host(host_actions[i % (W + K)])
    barrier()
    e2e_reps, e2e_times, t_all = 0, [], time.perf_counter()
    while e2e_reps < 3 or (time.perf_counter() - t_all < 0.15 and e2e_reps < 200):
        t0 = time.perf_counter()
        for i in range(W, W + K):
            obs, rew, done, info = env0.step_host(host_actions[i])
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
        e2e_reps += 1
    te = torch.tensor([sorted(e2e_times)[len(e2e_times) // 2]], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * E * K / float(te.item())
    checksum = float(np.asarray(rew, np.float64).sum())
    e2e_bytes = (env0.h2d_bytes_per_step(), env0.d2h_bytes_per_step(), env0.host_api_description(), env0.host_transport)
    # the other host transports of step_host, same loop (documents the trade-off; the headline e2e is the default one above)
    e2e_other = []
    for other in ("hybrid", "dense", "compact"):
        if other == e2e_bytes[3]:
            continue
        env0.host_transport = other
        for i in range(W):
            env0.step_host(host_actions[i])
        ot = []
        for _ in range(max(3, min(e2e_reps, 10))):
            t0 = time.perf_counter()
            for i in range(W, W + K):
                env0.step_host(host_actions[i])
            torch.cuda.synchronize()
            ot.append(time.perf_counter() - t0)
        to = torch.tensor([sorted(ot)[len(ot) // 2]], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(to, op=dist.ReduceOp.MAX)
        e2e_other.append({"host_transport": other, "value": world * E * K / float(to.item()), "d2h_bytes_per_step": env0.d2h_bytes_per_step() * world})
    env0.host_transport = e2e_bytes[3]

    for env in batches[1:]:
        env.buf = None
    batches = batches[:1]
    torch.cuda.empty_cache()
    dense = None if args.no_dense else dense_leg(args, dev, world, rank)
    iqn = None if args.no_iqn else iqn_bench(args, dev, world)       # every rank takes part (all-reduce inside)
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        abytes = algorithmic_bytes_per_env_step()
        achieved = E * abytes / (ms_per_step * 1e-3) / 1e9
        cpu_cores = os.cpu_count() or 1
        cfg = bench_config(E, world)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "timing": {"launch": f"timed region = {n_replay} replays of ONE CUDA graph of {n_launch} mnv_step launches ({reps} x steps={K}), "
                                 f"no eager launches; the launches alternate between {n_streams} stream(s) (independent env batches, "
                                 f"a batch stays on one stream); an event after every replay; ms_per_step = median replay / {n_launch}; "
                                 f"launch mode pdl={'2 (per call)' if params_pdl(params) else _lib.get_option('pdl')} "
                                 "(programmatic dependent launch, map tables fetched ahead of the grid dependency)",
                       "timed_region_ms": round(ms_region, 4), "ms_per_step_mean": ms_per_step_mean,
                       "ms_per_step_min": per_replay[0] / n_launch, "ms_per_step_max": per_replay[-1] / n_launch,
                       "mix_steps": args.mix, "streams": n_streams,
                       "single_stream": None if ms_1s is None else {
                           "ms_per_step": ms_1s, "roofline_frac": E * abytes / (ms_1s * 1e-3) / 1e9 / peak,
                           "what": "the same graph captured on ONE stream (every launch behind the grid dependency of the one before)"}},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH if E == ENVS_PER_GPU else None,
                         "traffic_source": NCU_TRAFFIC_SOURCE, "peak_source": peak_src, "algorithmic_bytes_per_env_step": abytes,
                         "kernel": "mnv_env_kernel<4,8,true,true>",
                         "launch_streams": n_streams,
                         "frac_single_stream": None if ms_1s is None else E * abytes / (ms_1s * 1e-3) / 1e9 / peak,
                         "note": f"achieved = algorithmic bytes per launch / period of back-to-back launches of independent env batches on {n_streams} "
                                 "stream(s); frac_single_stream = the same launches behind one another on ONE stream",
                         "issue_bound": {"warp_instructions_per_launch": NCU_WARP_INST_PER_LAUNCH,
                                         "floor_us_at_1_ipc_per_scheduler": NCU_WARP_INST_PER_LAUNCH / (148 * 4) / 1.965e3,
                                         "frac": NCU_WARP_INST_PER_LAUNCH / (148 * 4) / 1.965e3 / (ms_per_step * 1e3) if E == ENVS_PER_GPU else None,
                                         "note": "second roofline (SURVEY 8d): warp instructions of one launch (ncu smsp__inst_executed.sum) "
                                                 "/ (592 schedulers x 1.965 GHz) / measured time"}},
            "device_step_auto_reset": {"ms_per_step": ar_ms, "env_steps_per_s": world * E / (ar_ms * 1e-3),
                                       "finished_fraction_last_step": finished_frac,
                                       "what": "VecMarineNavEnv.step(auto_reset=True) device-resident: mnv_step + obs copy + masked mnv_reset + "
                                               "masked mnv_observe, CUDA graph of 16 steps, median of 12 replays"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e_bytes[0] * world,
                    "d2h_bytes_per_step": e2e_bytes[1] * world, "checksum": checksum,
                    "repetitions": e2e_reps, "api": e2e_bytes[2], "host_transport": e2e_bytes[3], "other_transport": e2e_other,
                    "auto_calibration_s_per_step": env0.host_transport_calibration},
            "gpu_launches": n_launch * n_replay,
            "clocks": clk.summary(),
        }
        if world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(E, cpu_cores)
        if dense is not None:
            line["dense"] = dense
        if iqn is not None:
            line["iqn"] = iqn
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
# Dense-map leg (BASELINE configs[4]: 32 obstacles, 64 beams, 131 072 envs over 8 GPUs = 16 384 per GPU; ray-cast stress)
# ---------------------------------------------------------------------------------------------------------------
def dense_leg(args, dev, world, rank):
    import torch
    import torch.distributed as dist
    from distributional_rl_navigation_b200 import env_ops
    from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv
    E, n_c, n_o, n_b, nb = 16384, 4, 32, 64, 8                      # 8 batches x 24 MB > L2
    batches = []
    for b in range(nb):
        env = VecMarineNavEnv(E, seed=(rank * nb + b) * E + 7_000_000, device=dev, num_cores=n_c, num_obs=n_o,
                              min_start_goal_dis=30.0, num_beams=n_b)
        env.reset()
        batches.append(env)
    g = torch.Generator(device=dev); g.manual_seed(99 + rank)
    actions = torch.randint(0, 9, (64, E), generator=g, device=dev, dtype=torch.int32)
    params = batches[0].params()
    for bi, env in enumerate(batches):
        for i in range(20):
            env.step(actions[(bi + i) % 64], auto_reset=True)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    n_streams = max(1, args.streams)                 # independent batches alternate between the capture streams (see run_b200)
    lanes = [side] + [torch.cuda.Stream(device=dev) for _ in range(n_streams - 1)]
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for s_ in lanes[1:]:
                s_.wait_stream(side)
            for i in range(64):
                with torch.cuda.stream(lanes[i % n_streams]):
                    env_ops.step(batches[i % nb].buf, params, action=actions[i])
            for s_ in lanes[1:]:
                side.wait_stream(s_)
    torch.cuda.current_stream().wait_stream(side)
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 256], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    ab = algorithmic_bytes_per_env_step(n_c, n_o, n_b)
    peak, _ = measured_peak_hbm()
    inst = 23.30e6          # smsp__inst_executed.sum of one launch, profiles/r2_dense_kernel_v4_ncu_full.csv
    floor_us = inst / (148 * 4) / 1.965e3
    return {"workload": "dense map: 32 obstacles, 64 sonar beams, 16384 envs/GPU (BASELINE configs[4] per-GPU shard), mnv_env_dense_kernel",
            "ms_per_step": ms, "env_steps_per_s": world * E / (ms * 1e-3), "streams": n_streams, "algorithmic_bytes_per_env_step": ab,
            "hbm_roofline_frac": E * ab / (ms * 1e-3) / 1e9 / peak, "bound": "issue (2048 ray-circle pairs per env-step), not HBM",
            "issue_bound": {"warp_instructions_per_launch": inst, "floor_us_at_1_ipc_per_scheduler": floor_us, "frac": floor_us / (ms * 1e3),
                            "note": "second roofline (SURVEY 8d): ncu smsp__inst_executed.sum of one launch / (592 schedulers x 1.965 GHz) / measured time"}}


# ---------------------------------------------------------------------------------------------------------------
# IQN legs (second half of the BASELINE metric: IQN updates/s; plus act and rollout+learn at BASELINE configs[2])
# ---------------------------------------------------------------------------------------------------------------
IQN_FLOP_PER_SAMPLE = 1813568          # fwd target + fwd local + bwd local, N = N' = 8
IQN_ACT_FLOP_PER_ENV = 2010816         # K = 32 forward


def iqn_bench(args, dev, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from distributional_rl_navigation_b200.iqn_agent import IQNAgent
    from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv
    from distributional_rl_navigation_b200 import iqn_ops

    out = {}
    B, E = 1024, args.envs
    agent = IQNAgent(26, 9, seed=0, device=dev, BATCH_SIZE=B)
    g = torch.Generator(device=dev); g.manual_seed(3 + int(os.environ.get("RANK", 0)))   # every replica trains on its own data
    n_sets = 64                                                            # 64 different batches (26 MB) rotated
    st = torch.randn(n_sets, B, 26, device=dev, generator=g) * 3; ns = torch.randn(n_sets, B, 26, device=dev, generator=g) * 3
    ac = torch.randint(0, 9, (n_sets, B), device=dev, generator=g); rw = torch.randn(n_sets, B, device=dev, generator=g)
    dn = (torch.rand(n_sets, B, device=dev, generator=g) < 0.05).float()
    tt = torch.rand(n_sets, B, 8, device=dev, generator=g); tl = torch.rand(n_sets, B, 8, device=dev, generator=g)

    def update(i):
        k = i % n_sets
        return agent.train_async((st[k], ac[k], rw[k], ns[k], dn[k]), (tt[k], tl[k]))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_upd = max(50, min(args.steps, 400))
    for i in range(10):
        update(i)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_upd):
        update(i)
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1) / n_upd], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_eager = float(ms.item())
    # the same updates as ONE CUDA graph per n_upd updates (IQNAgent.capture_updates: Adam's bias corrections from a device
    # control block): no Python between the launches -- with replicas, no rank is late for the gradient exchange
    replay = agent.capture_updates([(st[i % n_sets], ac[i % n_sets], rw[i % n_sets], ns[i % n_sets], dn[i % n_sets]) for i in range(n_upd)],
                                   [(tt[i % n_sets], tl[i % n_sets]) for i in range(n_upd)])
    replay(); sync()
    n_rep = 4
    e0.record()
    for _ in range(n_rep):
        replay()
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1) / (n_rep * n_upd)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    out["updates_per_s"] = 1e3 / ms
    out["update_ms"] = ms
    out["update_ms_eager_launches"] = ms_eager
    out["update_launch"] = f"{n_rep} replays of one CUDA graph of {n_upd} updates (IQNAgent.capture_updates); `update_ms_eager_launches` = the same updates launched from Python"
    peer = agent._tail is not None and agent._tail.world == world
    out["update_config"] = (f"batch {B} per GPU, N=N'=8, 3xTF32 mma.sync (fp32-level accuracy); two launches: iqn_loss_partials + iqn_update_tail "
                            "(tile-partial sum" + (", one-shot all-reduce of the 35785-float gradient over peer memory (NVLink)" if world > 1 and peer else "")
                            + ", clip_grad_norm_, Adam, kernel-side weight copies)") if agent.fused_tail and (world == 1 or peer) else \
                           (f"batch {B} per GPU, N=N'=8, 3xTF32 mma.sync; loss_grad + " + ("NCCL all-reduce(35785 f32) + " if world > 1 else "") + "clip_adam")
    out["update_tail_error_word"] = agent._tail.error() if agent._tail is not None else None
    out["update_gradient_exchange"] = "none (1 GPU)" if world == 1 else ("peer-memory one-shot all-reduce inside iqn_update_tail" if peer else "NCCL all_reduce")
    out["update_tflops_fp32"] = IQN_FLOP_PER_SAMPLE * B / (ms * 1e-3) / 1e12
    out["samples_per_s_all_gpus"] = world * B * 1e3 / ms
    out["loss_finite"] = bool(torch.isfinite(agent._loss).all().item())
    if world > 1:
        # data-parallel replicas: different batches per rank, one gradient all-reduce per update -> bit-identical weights
        chk = agent.qnetwork_local.flat.double().sum().reshape(1)
        hi, lo = chk.clone(), chk.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        pmax, pmin = agent.qnetwork_local.flat.clone(), agent.qnetwork_local.flat.clone()
        dist.all_reduce(pmax, op=dist.ReduceOp.MAX); dist.all_reduce(pmin, op=dist.ReduceOp.MIN)
        out["replicas_bit_identical"] = bool(torch.equal(pmax, pmin) and hi.item() == lo.item())

    # act: K = 32 forward + argmax for a whole env batch: tcgen05 kernel (bf16 operands) and the fp32 parity kernel
    obs = torch.randn(E, 26, device=dev, generator=g) * 3
    for name, tc, n_act in (("tc", True, 20), ("fp32", False, 3)):
        for _ in range(2):
            agent.act_batch(obs, 0.05, tensor_cores=tc)
        sync()
        e0.record()
        for _ in range(n_act):
            agent.act_batch(obs, 0.05, tensor_cores=tc)
        e1.record()
        sync()
        act_ms = e0.elapsed_time(e1) / n_act
        out[f"act_{name}_ms_per_env_batch"] = act_ms
        out[f"act_{name}_acts_per_s"] = world * E * 1e3 / act_ms
        out[f"act_{name}_tflops"] = IQN_ACT_FLOP_PER_ENV * E / (act_ms * 1e-3) / 1e12
    out["act_config"] = (f"{E} envs x K=32 taus, eps-greedy: tc = iqn_act_tc_sample (encoder pre-pass + tcgen05 kernel, bf16 operands, taus / "
                         "coin / random action from the in-kernel Philox stream); fp32 = torch.rand taus + iqn_forward + eager eps-greedy")

    # rollout + learn (BASELINE configs[2]): act -> env step (+auto-reset) -> replay append -> 1 update of 1024 per vector step
    env = VecMarineNavEnv(E, seed=12345 + E * (int(os.environ.get("RANK", 0))), device=dev, num_cores=N_CORES, num_obs=N_OBS,
                          min_start_goal_dis=30.0, num_beams=N_BEAMS)
    n_roll = 41                                                            # vector steps in the timed part (learn_vec's clock is GLOBAL transitions)
    for mode, tag in ((True, ""), (False, "_eager")):
        # graph=True: the vector step replayed as ONE CUDA graph (device control block for eps / counters / ring position);
        # graph=False: the same work launched from Python, ~12 launches per vector step (reported beside it)
        agent2 = IQNAgent(26, 9, seed=0, device=dev, BATCH_SIZE=B, BUFFER_SIZE=4 * E)
        n_warm = 6 if mode else 2
        agent2.learn_vec(total_timesteps=E * world * n_warm, train_env=env, batch_size=B, learning_starts=E, target_update_interval=100 * E * world,
                         graph=mode)
        sync()
        t0 = time.perf_counter()
        start_ts = agent2.current_timestep
        agent2.learn_vec(total_timesteps=start_ts + E * world * (n_roll - 1), train_env=env, batch_size=B, learning_starts=E,
                         target_update_interval=100 * E * world, graph=mode)
        sync()
        dt = time.perf_counter() - t0
        steps_done = agent2.current_timestep - start_ts
        out["rollout_learn" + tag + "_env_steps_per_s"] = steps_done / dt  # learn_vec counts GLOBAL transitions (E x world per vector step)
        out["rollout_learn" + tag + "_updates"] = agent2.optimizer.step_count
        del agent2
    out["rollout_learn_config"] = (f"{E} envs/GPU, eps-greedy IQN act K=32 (tcgen05), fused env step + auto-reset, device replay, 1 update of {B} per "
                                   "vector step; the vector step is ONE CUDA graph (learn_vec(graph=True)); `_eager` = launched from Python")

    if int(os.environ.get("RANK", 0)) == 0 and world == 1:
        # CPU baseline for the update: the numpy oracle (port of IQNAgent.train) on the host
        from oracle import iqn_oracle as io
        rs = np.random.RandomState(0)
        P = io.unflatten(agent.qnetwork_local.flat.cpu().numpy())
        b = 1024
        xs, x2 = rs.randn(b, 26).astype(np.float32), rs.randn(b, 26).astype(np.float32)
        a_, r_, d_ = rs.randint(0, 9, b), rs.randn(b).astype(np.float32), (rs.rand(b) < 0.05).astype(np.float32)
        t1, t2 = rs.rand(b, 8).astype(np.float32), rs.rand(b, 8).astype(np.float32)
        io.loss_and_grad(P, P, xs, a_, r_, x2, d_, t1, t2)
        t0 = time.perf_counter(); n = 0
        while time.perf_counter() - t0 < 5.0:
            io.loss_and_grad(P, P, xs, a_, r_, x2, d_, t1, t2); n += 1
        port_v = n / (time.perf_counter() - t0)
        out["cpu_baseline_updates_per_s"] = {"value": port_v, "kind": "port", "cores": os.cpu_count(),
                                             "sample": f"{n} updates of batch 1024, oracle/iqn_oracle.py (numpy/BLAS)"}
        from oracle import ref_timing
        if ref_timing.available():
            # the reference's own IQNAgent.train (thirdparty/IQN/agent.py:269-304, PyTorch CPU), staged in oracle/_ref
            r32, n32 = ref_timing.iqn_updates_per_s(32, threads=1, seconds=3.0)
            r1k, n1k = ref_timing.iqn_updates_per_s(1024, threads=1, seconds=4.0)
            r1k_all, n1k_all = ref_timing.iqn_updates_per_s(1024, threads=os.cpu_count() or 1, seconds=4.0)
            out["cpu_baseline_updates_per_s"] = {
                "value": max(r1k, r1k_all), "kind": "reference", "cores": os.cpu_count(),
                "batch_1024_1_thread": r1k, "batch_1024_all_threads": r1k_all, "batch_32_1_thread": r32,
                "sample": f"unmodified IQNAgent.train (PyTorch CPU): {n1k} / {n1k_all} updates of batch 1024 on 1 / all threads, {n32} updates of "
                          "batch 32 (the reference's default) on 1 thread",
                "port": {"value": port_v, "sample": f"{n} updates of batch 1024, oracle/iqn_oracle.py (numpy/BLAS)"}}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--streams", type=int, default=2, help="streams the rotating env batches' step launches alternate between (1, 2, 4 or 8)")
    ap.add_argument("--no-iqn", action="store_true", help="skip the IQN legs")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-map leg")
    ap.add_argument("--mix", type=int, default=200, help="auto-reset steps per env batch between the preload and the timed region")
    ap.add_argument("--preload", type=float, default=1.0, help="seconds of untimed identical load before the timed region (clock sampling)")
    args = ap.parse_args()
    claim_stdout()
    if args.impl == "reference":
        if args.steps > 50:
            args.steps = 50
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
