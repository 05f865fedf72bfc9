"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): the env kernels (sparse and dense shapes, auto-reset),
every host transport of step_host, the replay kernels and one captured rollout + learn vector step (control-block entry
points).  Small sizes: the tools slow kernels down by orders of magnitude."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv
for (nb, no, E) in ((11, 8, 1500), (64, 32, 300)):
    for transport in ("dense", "compact", "hybrid"):
        env = VecMarineNavEnv(E, seed=3, device="cuda:0", num_cores=4, num_obs=no, min_start_goal_dis=30.0, num_beams=nb, host_transport=transport)
        env.reset()
        g = torch.Generator(device="cuda"); g.manual_seed(0)
        for t in range(4):
            a = torch.randint(0, 9, (E,), device="cuda", generator=g, dtype=torch.int32)
            env.step(a)
        acts = np.random.RandomState(0).randint(0, 9, size=(3, E)).astype(np.int32)
        for t in range(3):
            env.step_host(acts[t])
        torch.cuda.synchronize()
if "--no-iqn" not in sys.argv:
    from distributional_rl_navigation_b200.iqn_agent import IQNAgent
    E, B = 512, 64
    env = VecMarineNavEnv(E, seed=5, device="cuda:0", num_cores=4, num_obs=8, min_start_goal_dis=30.0)
    agent = IQNAgent(26, 9, seed=0, device="cuda:0", BATCH_SIZE=B, BUFFER_SIZE=3 * E)
    agent.learn_vec(total_timesteps=E * 7, train_env=env, batch_size=B, learning_starts=E, target_update_interval=2 * E, graph=True,
                    sample_without_replacement=True)
    torch.cuda.synchronize()
print("done")
