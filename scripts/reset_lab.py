#!/usr/bin/env python
"""Lab: device-side cost of the auto-reset path of VecMarineNavEnv.step at a given fraction of finished environments
(CUDA events around graph replays of each component).   python scripts/reset_lab.py [E] [fractions...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distributional_rl_navigation_b200 import env_ops
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
fracs = [float(x) for x in sys.argv[2:]] or [0.0016, 0.01, 0.05]
env = VecMarineNavEnv(E, seed=0, device="cuda:0", num_cores=4, num_obs=8, min_start_goal_dis=30.0)
env.reset()
b, p, rp = env.buf, env.params(), env.reset_params()
g = torch.Generator(device="cuda:0"); g.manual_seed(0)


def timed(fn, n=20):
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
        with torch.cuda.graph(gr, stream=side):
            for _ in range(n):
                fn()
    torch.cuda.current_stream().wait_stream(side)
    gr.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / n)
    return best


print(f"E={E}: obs copy {timed(lambda: b['obs'].copy_(b['next_obs'])):.2f} us, step {timed(lambda: env_ops.step(b, p, obs=b['next_obs'])):.2f} us")
for f in fracs:
    mask = (torch.rand(E, device="cuda:0", generator=g) < f).to(torch.uint8)
    t_reset = timed(lambda: env_ops.reset(b, env.rng_key, env.rng_pos, rp, mask=mask))
    t_obs = timed(lambda: env_ops.observe(b, p, mask=mask, velocity_from_state=True))
    print(f"  finished fraction {f:.4f} ({int(mask.sum())} envs): masked reset {t_reset:.2f} us, masked observe {t_obs:.2f} us", flush=True)
