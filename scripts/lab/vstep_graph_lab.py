"""lab: vector-step period of IQNAgent.learn_vec, eager launches vs the captured vector step (graph=True), by batch of envs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from distributional_rl_navigation_b200.iqn_agent import IQNAgent
from distributional_rl_navigation_b200.vec_env import VecMarineNavEnv

for E, B, U in ((64, 32, 16), (1024, 256, 1), (4096, 1024, 1), (16384, 1024, 1), (65536, 1024, 1)):
    for mode in (False, True):
        env = VecMarineNavEnv(E, seed=1, device="cuda:0", num_cores=4, num_obs=8, min_start_goal_dis=30.0)
        agent = IQNAgent(26, 9, seed=0, device="cuda:0", BATCH_SIZE=B, BUFFER_SIZE=max(4 * E, 8 * B))
        kw = dict(train_env=env, batch_size=B, learning_starts=2 * B, target_update_interval=10 ** 9, updates_per_step=U, graph=mode)
        agent.learn_vec(total_timesteps=E * max(8, 4 * B // E + 8), **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); s0 = agent.current_timestep
        agent.learn_vec(total_timesteps=s0 + E * 200, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = (agent.current_timestep - s0) / E
        print(f"E={E:6d} B={B:5d} updates/step={U:2d} graph={mode!s:5}: {dt / n * 1e6:8.1f} us per vector step, "
              f"{(agent.current_timestep - s0) / dt / 1e6:8.2f} M env-steps/s, {agent.optimizer.step_count} updates", flush=True)
        del env, agent
        torch.cuda.empty_cache()
