"""Module path of the reference robot model (marinenav_env/envs/utils/robot.py)."""
from distributional_rl_navigation_b200.marinenav_env import Robot, Sonar  # noqa: F401
