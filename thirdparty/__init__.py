"""Drop-in for `from thirdparty import IQNAgent` (train_IQN_model.py:3): the B200 IQN agent."""
from distributional_rl_navigation_b200.iqn_agent import IQNAgent  # noqa: F401
