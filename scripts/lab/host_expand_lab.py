"""lab: the host transports of step_host side by side -- dense, compact (by expander threads), hybrid (by compact share)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from distributional_rl_navigation_b200 import _hostlib
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv

E = 65536
print("cpus:", len(os.sched_getaffinity(0)), flush=True)
acts = np.random.RandomState(0).randint(0, 9, size=(64, E)).astype(np.int32)
cases = [("dense", 0, None)] + [("compact", nt, None) for nt in (4, 8)] + \
        [("hybrid", nt, f) for nt in (4, 8) for f in ("0.25", "0.375", "0.5", "0.625", "0.75")] + [("dense", 0, None)]
for transport, nt, frac in cases:
    if frac is not None:
        os.environ["MNV_HOST_HYBRID_FRACTION"] = frac
    if nt:
        os.environ["MNV_HOST_THREADS"] = str(nt)
    env = VecMarineNavEnv(E, seed=0, device="cuda:0", num_cores=4, num_obs=8, min_start_goal_dis=30.0, host_transport=transport)
    env.reset_host()
    for i in range(10):
        env.step_host(acts[i])
    best = 1e9
    for rep in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 100
        for i in range(n):
            env.step_host(acts[i % 64])
        torch.cuda.synchronize(); best = min(best, (time.perf_counter() - t0) / n)
    print(f"{transport:8s} threads={nt:2d} compact share={frac or '-':>5}: step_host {best * 1e6:7.1f} us  {E / best / 1e6:7.1f} M env-steps/s", flush=True)
    del env
