"""TEST / BENCH INFRASTRUCTURE ONLY -- times the reference's OWN Python implementation of the hot path on host cores.

bench.py's CPU-baseline legs call this (never the product).  The reference steps ONE environment per process
(marinenav_env.py:199), so "all host cores" = P independent worker processes, one MarineNavEnv each, disjoint seeds
(CPU reference timing).  Workers are separate interpreters (``python -m oracle.ref_timing
--worker env ...``): they never share the parent's CUDA context and each reports the durations of its own timed segments.

Workload of a worker = bench.py's: maps from the reference's reset rules with fixed counts (4 cores / 8 obstacles, start-goal
distance > 30, 11 beams), uniform random actions, reset on done (the reference's reset prints are swallowed).
"""
import contextlib
import io
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def available():
    from oracle import ref_import
    return ref_import.reference_available()


# ---------------------------------------------------------------------------------------------------------------------
# workers (run in their own interpreter)
# ---------------------------------------------------------------------------------------------------------------------
def _worker_env(seed, n_segments, steps_per_segment, n_c, n_o, n_b, warm):
    import numpy as np
    from oracle import ref_import
    ref_env, _, _ = ref_import.load_reference()
    sink = io.StringIO()
    env = ref_env.MarineNavEnv(seed=seed)
    env.num_cores, env.num_obs, env.min_start_goal_dis = n_c, n_o, 30.0
    if n_b != env.robot.sonar.num_beams:                       # dense maps: 64 beams (robot.py:7-21)
        env.robot.sonar.num_beams = n_b
        env.robot.sonar.compute_phi(); env.robot.sonar.compute_beam_angles()
    with contextlib.redirect_stdout(sink):
        env.reset()
    rs = np.random.RandomState(123 + seed)

    def run(n):
        for _ in range(n):
            _, _, done, _ = env.step(int(rs.randint(9)))
            if done:
                sink.seek(0); sink.truncate()
                with contextlib.redirect_stdout(sink):
                    env.reset()
    run(warm)
    durs = []
    for _ in range(n_segments):
        t0 = time.perf_counter()
        run(steps_per_segment)
        durs.append(time.perf_counter() - t0)
    return durs


def _worker_iqn(batch, threads, seconds):
    import numpy as np
    import torch
    from oracle import ref_import
    torch.set_num_threads(threads)
    _, ref_agent, _ = ref_import.load_reference()
    agent = ref_agent.IQNAgent(26, 9, BATCH_SIZE=batch, seed=0)
    rs = np.random.RandomState(0)
    t = lambda a: torch.from_numpy(a)
    exp = (t(rs.randn(batch, 26).astype(np.float32) * 3), t(rs.randint(0, 9, (batch, 1)).astype(np.int64)),
           t(rs.randn(batch, 1).astype(np.float32)), t(rs.randn(batch, 26).astype(np.float32) * 3),
           t((rs.rand(batch, 1) < 0.05).astype(np.float32)))
    for _ in range(3):
        agent.train(exp)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        agent.train(exp); n += 1
    return [n, time.perf_counter() - t0]


# ---------------------------------------------------------------------------------------------------------------------
# parent side
# ---------------------------------------------------------------------------------------------------------------------
def _spawn(args):
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    env["CUDA_VISIBLE_DEVICES"] = ""                         # the baseline is a CPU run
    env.setdefault("OMP_NUM_THREADS", "1")
    return subprocess.Popen([sys.executable, "-m", "oracle.ref_timing", "--worker"] + [str(a) for a in args], cwd=ROOT, env=env,
                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)


def env_segments(n_procs, n_segments, steps_per_segment, n_c=4, n_o=8, n_b=11, warm=20, seed0=0):
    """P worker processes x n_segments timed segments of steps_per_segment env steps each.
    Returns durs[p][i] (seconds).  Aggregate rate of segment i = sum_p steps_per_segment / durs[p][i]."""
    procs = [_spawn(["env", seed0 + p, n_segments, steps_per_segment, n_c, n_o, n_b, warm]) for p in range(n_procs)]
    out = []
    for p in procs:
        txt, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("reference worker failed")
        out.append(json.loads(txt.strip().splitlines()[-1]))
    return out


def env_steps_per_s(n_procs, steps_per_proc, **kw):
    durs = env_segments(n_procs, 1, steps_per_proc, **kw)
    return sum(steps_per_proc / d[0] for d in durs), max(d[0] for d in durs)


def iqn_updates_per_s(batch, threads=1, seconds=4.0):
    p = _spawn(["iqn", batch, threads, seconds])
    txt, _ = p.communicate()
    if p.returncode != 0:
        raise RuntimeError("reference IQN worker failed")
    n, dt = json.loads(txt.strip().splitlines()[-1])
    return n / dt, n


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--worker":
        kind, a = sys.argv[2], sys.argv[3:]
        real = sys.stdout
        sys.stdout = io.StringIO()                            # keep the reference's own prints off the result channel
        if kind == "env":
            res = _worker_env(int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(a[4]), int(a[5]), int(a[6]))
        else:
            res = _worker_iqn(int(a[0]), int(a[1]), float(a[2]))
        real.write(json.dumps(res) + "\n")
        real.flush()
