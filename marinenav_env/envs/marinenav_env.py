"""Module path of the reference env (`import marinenav_env.envs.marinenav_env as marinenav_env`, run_experiments.py:4)."""
from distributional_rl_navigation_b200.marinenav_env import Core, MarineNavEnv, Obstacle  # noqa: F401
