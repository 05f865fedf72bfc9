"""Device-time of the fused step kernel on the BASELINE configs that are not the bench line (parity-tested elsewhere).
    python scripts/measure_configs.py        (needs a GPU)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from distributional_rl_navigation_b200 import env_ops  # noqa: E402
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv  # noqa: E402


def measure(E, nc, no, nb, n_batches=4, steps=200):
    envs = []
    for b in range(n_batches):
        env = VecMarineNavEnv(E, seed=b * E, device="cuda:0", num_cores=nc, num_obs=no, min_start_goal_dis=30.0, num_beams=nb)
        env.reset()
        env.rng_key = None
        envs.append(env)
    p = envs[0].params()
    act = torch.randint(0, 9, (E,), device="cuda", dtype=torch.int32)
    for i in range(8):
        env_ops.step(envs[i % n_batches].buf, p, action=act)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.cuda.graph(g, stream=s):
        for i in range(n_batches * 4):
            env_ops.step(envs[i % n_batches].buf, p, action=act)
    torch.cuda.current_stream().wait_stream(s)
    reps = max(1, steps // (n_batches * 4))
    g.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (reps * n_batches * 4)
    bytes_per = 8 * (4 + 2 + 3 * no + 3 * nc) + 8 + (8 * 4 + 4 + 4 * (4 + 2 * nb) + 4 + 2)
    print(f"E={E:7d} cores={nc} obstacles={no:2d} beams={nb:2d}: {us:8.2f} us/step  {E / us:9.1f} M env-steps/s  "
          f"{E * bytes_per / us / 1e3:7.1f} GB/s algorithmic ({bytes_per} B/env-step)")


if __name__ == "__main__":
    measure(65536, 4, 8, 11, n_batches=8)        # BASELINE configs[1] (the bench line)
    measure(65536, 8, 10, 11, n_batches=8)       # last curriculum stage (8 cores, 10 obstacles)
    measure(16384, 4, 32, 64, n_batches=8)       # BASELINE configs[4] per GPU: dense ray-cast stress
    measure(1, 4, 8, 11, n_batches=1)            # single-env facade launch
