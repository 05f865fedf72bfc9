from distributional_rl_navigation_b200.iqn_agent import IQNAgent  # noqa: F401
