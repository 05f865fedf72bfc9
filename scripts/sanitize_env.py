import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv
for (nb, no, E) in ((11, 8, 1500), (64, 32, 300)):
    env = VecMarineNavEnv(E, seed=3, device="cuda:0", num_cores=4, num_obs=no, min_start_goal_dis=30.0, num_beams=nb)
    env.reset()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for t in range(6):
        a = torch.randint(0, 9, (E,), device="cuda", generator=g, dtype=torch.int32)
        env.step(a)
    acts = np.random.RandomState(0).randint(0, 9, size=(3, E)).astype(np.int32)
    for t in range(3):
        env.step_host(acts[t])
    torch.cuda.synchronize()
print("done")
