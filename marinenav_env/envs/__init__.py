from distributional_rl_navigation_b200.marinenav_env import MarineNavEnv, Core, Obstacle, Robot, Sonar  # noqa: F401
