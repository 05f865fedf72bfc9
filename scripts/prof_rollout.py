"""Launch list of the rollout + learn loop (BASELINE configs[2]): run under
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rollout_launches.csv python scripts/prof_rollout.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributional_rl_navigation_b200.iqn_agent import IQNAgent
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv

E, B = 65536, 1024
env = VecMarineNavEnv(E, seed=1, device="cuda:0", num_cores=4, num_obs=8, min_start_goal_dis=30.0)
agent = IQNAgent(26, 9, seed=0, device="cuda:0", BATCH_SIZE=B, BUFFER_SIZE=4 * E)
agent.learn_vec(total_timesteps=E * 2, train_env=env, batch_size=B, learning_starts=E, target_update_interval=100 * E)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
agent.learn_vec(total_timesteps=agent.current_timestep + E * 5, train_env=env, batch_size=B, learning_starts=E, target_update_interval=100 * E)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
