from distributional_rl_navigation_b200.replay_buffer import ReplayBuffer  # noqa: F401
