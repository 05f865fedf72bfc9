"""lab: the compact host transport of step_host taken apart -- graph replay + sync vs the native expander, by thread count."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from distributional_rl_navigation_b200 import _hostlib
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv

E = 65536
print("cpus:", len(os.sched_getaffinity(0)), flush=True)
acts = np.random.RandomState(0).randint(0, 9, size=(64, E)).astype(np.int32)
for transport in ("dense", "compact"):
    for nt in ((0,) if transport == "dense" else (1, 2, 4, 8, 12, 16)):
        env = VecMarineNavEnv(E, seed=0, device="cuda:0", num_cores=4, num_obs=8, min_start_goal_dis=30.0, host_transport=transport)
        env.reset()
        pin = env._pin()
        if nt:
            env._expander = _hostlib.Expander(E, env.obs_dim, n_threads=nt)
        t_exp = [0.0]
        if transport == "compact":
            real = env._expander.expand
            def timed(*a, _r=real):
                t0 = time.perf_counter(); _r(*a); t_exp[0] += time.perf_counter() - t0
            env._expander.expand = timed
        for i in range(10):
            env.step_host(acts[i])
        t_exp[0] = 0.0
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 200
        for i in range(n):
            env.step_host(acts[i % 64])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{transport:8s} threads={nt:2d}: step_host {dt / n * 1e6:7.1f} us, of which expander {t_exp[0] / n * 1e6:7.1f} us", flush=True)
        del env
