"""ncu driver for the fused step kernel in its bench state (65 536 envs, stationary rollout distribution):
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mnv_env_kernel -c 2 -o gpurun_out/step python scripts/prof_step.py
The profiled launches are the plain mnv_step launches between cudaProfilerStart / Stop."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distributional_rl_navigation_b200 import env_ops
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dense = len(sys.argv) > 2 and sys.argv[2] == "dense"
n_o, n_b = (32, 64) if dense else (8, 11)
envs = [VecMarineNavEnv(E, seed=b * E, device="cuda:0", num_cores=4, num_obs=n_o, min_start_goal_dis=30.0, num_beams=n_b) for b in range(4)]
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
actions = torch.randint(0, 9, (64, E), generator=g, device="cuda:0", dtype=torch.int32)
for bi, env in enumerate(envs):
    env.reset()
    for i in range(100):
        env.step(actions[(bi + i) % 64], auto_reset=True)
params = envs[0].params()
for i in range(8):
    env_ops.step(envs[i % 4].buf, params, action=actions[i])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(4):
    env_ops.step(envs[i % 4].buf, params, action=actions[8 + i])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
