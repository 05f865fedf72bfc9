#!/usr/bin/env python
"""Run an UNMODIFIED driver script of the reference (train_IQN_model.py, ...) on the B200 drop-in packages.

    python scripts/run_reference_driver.py /path/to/reference/train_IQN_model.py -C config/config_IQN.json -D cuda:0

Why a launcher: `python /path/to/reference/train_IQN_model.py` puts the SCRIPT's directory at sys.path[0], ahead of
PYTHONPATH, so `marinenav_env` / `thirdparty` would resolve to the reference's own packages, and the script itself runs
`sys.path.insert(0, "./thirdparty")` (train_IQN_model.py:1-2).  This launcher
  1. puts THIS repository's root first on sys.path and imports its `marinenav_env`, `marinenav_env.envs.marinenav_env` and
     `thirdparty` packages up front -- later path edits by the driver cannot re-resolve names that are already imported;
  2. installs distributional_rl_navigation_b200.compat_gym as `gym` when the real gym package is absent
     (train_IQN_model.py:4 `import gym`, :96,:100 `gym.make('marinenav_env:marinenav_env-v0', ...)`);
  3. executes the driver file as __main__ with runpy (which, unlike `python file.py`, does not add the file's directory
     to sys.path).  The driver's source is not touched.
"""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN_PACKAGES = ("marinenav_env", "marinenav_env.envs", "marinenav_env.envs.marinenav_env", "marinenav_env.envs.utils.robot",
                   "thirdparty", "thirdparty.IQN", "thirdparty.IQN.agent", "thirdparty.IQN.model", "thirdparty.IQN.replay_buffer")


def install_gym():
    """`import gym` -> the real package if installed, else the bundled stand-in (Env, spaces, register, make)."""
    try:
        import gym  # noqa: F401
        return sys.modules["gym"]
    except ImportError:
        import types
        from distributional_rl_navigation_b200 import compat_gym
        sys.modules["gym"] = compat_gym
        spaces = types.ModuleType("gym.spaces")
        spaces.Discrete, spaces.Box = compat_gym.Discrete, compat_gym.Box
        envs = types.ModuleType("gym.envs")
        registration = types.ModuleType("gym.envs.registration")
        registration.register = compat_gym.register
        envs.registration = registration
        sys.modules.update({"gym.spaces": spaces, "gym.envs": envs, "gym.envs.registration": registration})
        return compat_gym


def prepare():
    """Steps 1 + 2; returns the module objects of the drop-in packages (all inside this repository)."""
    import importlib
    sys.path[:] = [ROOT] + [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    for name in [m for m in sys.modules if m.split(".")[0] in ("marinenav_env", "thirdparty", "IQN")]:
        f = getattr(sys.modules[name], "__file__", None) or ""
        if not os.path.abspath(f).startswith(ROOT + os.sep):
            del sys.modules[name]                              # something else already claimed the name: re-resolve it here
    install_gym()
    mods = {name: importlib.import_module(name) for name in DROPIN_PACKAGES}
    for name, mod in mods.items():
        f = os.path.abspath(getattr(mod, "__file__", "") or "")
        if not f.startswith(ROOT + os.sep):
            raise ImportError(f"{name} resolved to {f}, not to the drop-in package under {ROOT}")
    return mods


def run(driver, argv=(), run_name="__main__"):
    prepare()
    driver = os.path.abspath(driver)
    old_argv = sys.argv
    sys.argv = [driver] + list(argv)
    try:
        return runpy.run_path(driver, run_name=run_name)
    finally:
        sys.argv = old_argv


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    run(sys.argv[1], sys.argv[2:])
