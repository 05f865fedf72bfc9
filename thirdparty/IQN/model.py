from distributional_rl_navigation_b200.iqn_model import ObsEncoder  # noqa: F401
