"""Where the end-to-end (host buffers) step time goes: python scripts/e2e_breakdown.py (needs a GPU)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distributional_rl_navigation_b200 import env_ops
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv

E = 65536
env = VecMarineNavEnv(E, seed=0, device="cuda:0", num_cores=4, num_obs=8, min_start_goal_dis=30.0)
env.reset()
acts = np.random.RandomState(0).randint(0, 9, size=(60, E)).astype(np.int32)
for i in range(10):
    env.step_host(acts[i])


def timeit(fn, n=40):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t0) / n


pin = env._pin()
a_dev = torch.from_numpy(acts[0]).cuda()
print("step_host (CUDA graph)     %8.1f us" % timeit(lambda i: env.step_host(acts[10 + i % 50])))
print("step_host (eager)          %8.1f us" % timeit(lambda i: env.step_host(acts[10 + i % 50], graph=False)))
print("device step(auto_reset)    %8.1f us" % timeit(lambda i: env.step(a_dev, auto_reset=True)))
print("device step(no reset)      %8.1f us" % timeit(lambda i: env.step(a_dev, auto_reset=False)))
print("H2D actions (pinned)       %8.1f us" % timeit(lambda i: env.buf["action"].copy_(pin["action"], non_blocking=True)))
def d2h(i):
    pin["obs"].copy_(env.buf["obs"], non_blocking=True); pin["reward"].copy_(env.buf["reward"], non_blocking=True)
    pin["done"].copy_(env.buf["done"], non_blocking=True); pin["info"].copy_(env.buf["info"], non_blocking=True)
    torch.cuda.current_stream().synchronize()
print("D2H obs+reward+done+info   %8.1f us  (%.1f GB/s)" % ((t := timeit(d2h)), env.d2h_bytes_per_step() / t / 1e3))
print("host copy of actions       %8.1f us" % timeit(lambda i: pin["action"].copy_(torch.as_tensor(acts[i % 50], dtype=torch.int32))))
print("done envs per step: %.1f" % float(env.buf["done"].float().sum().item()))
