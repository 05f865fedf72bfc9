"""Single-env drop-in loop rate (MarineNavEnv facade + IQNAgent.act), the path train_IQN_model.py exercises unchanged."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import marinenav_env  # noqa: F401
from distributional_rl_navigation_b200 import marinenav_env as impl
from thirdparty import IQNAgent

env = impl._gym.make('marinenav_env:marinenav_env-v0', seed=0)
env.num_cores, env.num_obs = 4, 8
agent = IQNAgent(26, 9, seed=0, device="cuda:0")
obs = env.reset()
rng = np.random.RandomState(0)
for phase, n in (("env.step only", 2000), ("act + env.step", 2000)):
    t0 = time.perf_counter()
    for i in range(n):
        a = int(rng.randint(9)) if phase == "env.step only" else int(agent.act(obs, 0.05))
        obs, r, done, info = env.step(a)
        if done:
            obs = env.reset()
    dt = time.perf_counter() - t0
    print(f"{phase:16s}: {n / dt:8.0f} steps/s ({1e6 * dt / n:6.1f} us/step)   [reference, 1 CPU core: 330-385 steps/s env only]")
