"""lab: vector-step period of IQNAgent.learn_vec (65 536 envs) for policy_lag 0 / 1, with and without updates."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from distributional_rl_navigation_b200.iqn_agent import IQNAgent
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv

E, B = 65536, 1024
for lag in (0, 1):
    for learn in (True, False):
        env = VecMarineNavEnv(E, seed=1, device="cuda:0", num_cores=4, num_obs=8, min_start_goal_dis=30.0)
        agent = IQNAgent(26, 9, seed=0, device="cuda:0", BATCH_SIZE=B, BUFFER_SIZE=4 * E)
        ls = E if learn else 10 ** 12
        agent.learn_vec(total_timesteps=E * 3, train_env=env, batch_size=B, learning_starts=ls, target_update_interval=100 * E, policy_lag=lag)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); s0 = agent.current_timestep
        agent.learn_vec(total_timesteps=s0 + E * 40, train_env=env, batch_size=B, learning_starts=ls, target_update_interval=100 * E, policy_lag=lag)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = (agent.current_timestep - s0) / E
        print(f"policy_lag={lag} updates={learn}: {dt / n * 1e6:.1f} us per vector step, {(agent.current_timestep - s0) / dt / 1e6:.1f} M env-steps/s", flush=True)
        del env, agent
        torch.cuda.empty_cache()
