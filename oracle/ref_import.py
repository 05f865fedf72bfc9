"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference in-process.

Used by the fixture generators under ``tests/golden/`` and by the oracle-pinning
tests (skipped when the reference tree is absent, e.g. on the GPU box).  Nothing
in the product package may import this module.

The reference (RobustFieldAutonomyLab/Distributional_RL_Navigation @ e77bbbf)
needs three shims to import under numpy 2.x without ``gym`` installed
(SURVEY.md section 8(c)); no reference file is modified:

1. a ``gym`` stub providing ``Env``, ``spaces.Discrete``, ``spaces.Box`` and
   ``envs.registration.register`` (used at marinenav_env.py:4,25,33,35 and
   marinenav_env/__init__.py:1);
2. ``np.infty`` (robot.py:148), removed in numpy 2.0;
3. a ``thirdparty`` package stub so that ``thirdparty/__init__.py:1-4`` (which
   drags in stable-baselines3 -> gym) is not executed.
"""
import os
import sys
import types
import warnings

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/stage_ref.py (travels to the GPU box)


def _find_root():
    cands = [os.environ.get("MARINENAV_REF"), "/root/reference", _STAGED]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "marinenav_env", "envs", "marinenav_env.py")):
            return c
    return "/root/reference"


REF_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "marinenav_env", "envs", "marinenav_env.py"))


def _install_gym_stub():
    if "gym" in sys.modules and not getattr(sys.modules["gym"], "_mnv_oracle_stub", False):
        return
    gym = types.ModuleType("gym")
    gym._mnv_oracle_stub = True

    class Env:  # noqa: D401 - minimal stand-in for gym.Env
        def close(self):
            pass

    class Discrete:
        def __init__(self, n):
            self.n = int(n)

    class Box:
        def __init__(self, low, high, dtype=None):
            self.low, self.high, self.dtype = low, high, dtype
            self.shape = getattr(low, "shape", None)

    spaces = types.ModuleType("gym.spaces")
    spaces.Discrete, spaces.Box = Discrete, Box
    envs = types.ModuleType("gym.envs")
    registration = types.ModuleType("gym.envs.registration")
    registration.register = lambda **kw: None
    envs.registration = registration
    gym.Env, gym.spaces, gym.envs = Env, spaces, envs
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.envs": envs,
                        "gym.envs.registration": registration})


def load_reference():
    """Return (marinenav_env_module, IQN agent module, IQN model module) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import numpy as np
    if not hasattr(np, "infty"):
        np.infty = np.inf
    warnings.filterwarnings("ignore", category=PendingDeprecationWarning)
    _install_gym_stub()
    # our repo root also has drop-in ``marinenav_env``/``thirdparty`` packages:
    # make sure the names resolve to the REFERENCE here.
    for name in [m for m in sys.modules if m == "marinenav_env" or m.startswith("marinenav_env.")
                 or m == "thirdparty" or m.startswith("thirdparty.") or m == "IQN" or m.startswith("IQN.")]:
        mod = sys.modules[name]
        f = getattr(mod, "__file__", None) or ""
        if not f.startswith(REF_ROOT) and not getattr(mod, "_mnv_oracle_stub", False):
            del sys.modules[name]
    tp = types.ModuleType("thirdparty")
    tp.__path__ = [os.path.join(REF_ROOT, "thirdparty")]
    tp._mnv_oracle_stub = True
    sys.modules["thirdparty"] = tp
    saved = list(sys.path)
    try:
        sys.path.insert(0, os.path.join(REF_ROOT, "thirdparty"))
        sys.path.insert(0, REF_ROOT)
        import marinenav_env.envs.marinenav_env as ref_env
        import thirdparty.IQN.agent as ref_agent
        import thirdparty.IQN.model as ref_model
    finally:
        sys.path[:] = saved
    return ref_env, ref_agent, ref_model
