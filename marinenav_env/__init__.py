"""Drop-in package name of the reference (marinenav_env/__init__.py:1-5): registers 'marinenav_env-v0' so that
gym.make('marinenav_env:marinenav_env-v0', seed=..., schedule=...) resolves to the B200 implementation."""
try:
    from gym.envs.registration import register
except ImportError:                                     # gym is not installed: the bundled stand-in registry
    from distributional_rl_navigation_b200.compat_gym import register

register(
    id='marinenav_env-v0',
    entry_point='marinenav_env.envs:MarineNavEnv',
)
