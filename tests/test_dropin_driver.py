"""The reference's training driver runs UNMODIFIED on the drop-in packages (BASELINE north_star: "drops into
train_IQN_model.py unchanged"; train_IQN_model.py:1-13,96-121).

The driver file is the reference's own, either in /root/reference or staged (unmodified, git-ignored) in oracle/_ref by
oracle/stage_ref.py; scripts/run_reference_driver.py executes it with runpy after putting this repository's packages in
front.  Every check runs in a fresh interpreter so that the test process's own imports play no part.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAUNCHER = os.path.join(ROOT, "scripts", "run_reference_driver.py")


def _driver():
    for base in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        f = os.path.join(base, "train_IQN_model.py")
        if os.path.isfile(f):
            return f
    return None


needs_driver = pytest.mark.skipif(_driver() is None, reason="reference driver not available (neither /root/reference nor oracle/_ref)")

_RESOLVE = r"""
import json, os, sys
sys.path.insert(0, os.path.join({root!r}, "scripts"))
import run_reference_driver as launcher
g = launcher.run({driver!r}, [], run_name="imported_not_main")      # module level of the driver: its imports and defs
import importlib
out = dict(
    iqn_module=g["IQNAgent"].__module__,
    env_module_file=os.path.abspath(g["marinenav_env"].__file__),
    thirdparty_file=os.path.abspath(sys.modules["thirdparty"].__file__),
    marinenav_file=os.path.abspath(sys.modules["marinenav_env"].__file__),
    gym_name=g["gym"].__name__,
    entry_is_ours=importlib.import_module("marinenav_env.envs").MarineNavEnv.__module__,
    driver_path_edit="./thirdparty" in sys.path,
)
try:
    env = g["gym"].make("marinenav_env:marinenav_env-v0", seed=3)
    out["made"] = type(env).__module__ + "." + type(env).__name__
except Exception as e:
    out["make_error"] = type(e).__module__ + "." + type(e).__name__
print("RESULT" + json.dumps(out))
"""


@needs_driver
@pytest.mark.reference
def test_unmodified_driver_resolves_to_the_dropin_packages():
    """Module level of the unmodified driver: `from thirdparty import IQNAgent`, `import gym`, `import
    marinenav_env.envs.marinenav_env` land in THIS repository although the driver prepends ./thirdparty to sys.path and lives
    next to the reference's own packages; gym.make reaches this repository's MarineNavEnv (which, without a GPU, fails loudly)."""
    code = _RESOLVE.format(root=ROOT, driver=_driver())
    cwd = os.path.dirname(_driver())                            # the worst case: cwd = the reference tree itself
    r = subprocess.run([sys.executable, "-c", code], cwd=cwd, capture_output=True, text=True, timeout=300,
                       env={**os.environ, "CUDA_VISIBLE_DEVICES": os.environ.get("CUDA_VISIBLE_DEVICES", "")})
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][-1][6:])
    assert out["iqn_module"] == "distributional_rl_navigation_b200.iqn_agent"
    assert out["env_module_file"] == os.path.join(ROOT, "marinenav_env", "envs", "marinenav_env.py")
    assert out["thirdparty_file"] == os.path.join(ROOT, "thirdparty", "__init__.py")
    assert out["marinenav_file"] == os.path.join(ROOT, "marinenav_env", "__init__.py")
    assert out["entry_is_ours"] == "distributional_rl_navigation_b200.marinenav_env"
    assert out["driver_path_edit"], "the driver's own sys.path.insert(0, './thirdparty') ran (the file is unmodified)"
    import torch
    if torch.cuda.is_available():
        assert out.get("made") == "distributional_rl_navigation_b200.marinenav_env.MarineNavEnv", out
    else:
        assert out.get("make_error") == "distributional_rl_navigation_b200._lib.MarinenavError", out   # no CPU fallback


@needs_driver
@pytest.mark.gpu
def test_unmodified_driver_trains_end_to_end(tmp_path, golden_dir):
    """python scripts/run_reference_driver.py train_IQN_model.py -C cfg -D cuda:0 : 10 200 single-env steps through
    gym.make / IQNAgent.learn (agent.py:94-173): learning starts at 10 000, so the run includes the 30-map greedy + adaptive
    evaluations, 50 IQN updates and the checkpoint.  The eval_config.json the DRIVER writes (create_eval_configs,
    train_IQN_model.py:123-148, seed 348) must equal the reference's shipped file bit for bit."""
    cfg = tmp_path / "cfg.json"
    cfg.write_text(json.dumps({"agent": "IQN", "seed": [7], "total_timesteps": 10200, "eval_freq": 100000,
                               "save_dir": str(tmp_path / "out")}))
    r = subprocess.run([sys.executable, LAUNCHER, _driver(), "-C", str(cfg), "-D", "cuda:0"], cwd=str(tmp_path),
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    runs = list((tmp_path / "out").glob("training_*/seed_7"))
    assert len(runs) == 1
    d = runs[0]
    for f in ("trial_config.json", "training_schedule.json", "eval_config.json", "greedy_evaluations.npz",
              "adaptive_evaluations.npz", "network_params.pth", "constructor_params.json"):
        assert (d / f).is_file(), f
    with open(d / "eval_config.json") as f, open(os.path.join(golden_dir, "eval_config.json")) as g:
        mine, shipped = json.load(f), json.load(g)
    for c in mine.values():                 # episode_data() of the current reference also records the (empty) trajectory
        assert c["robot"].pop("trajectory") == []     # (marinenav_env.py:620); the shipped file predates that key
    assert mine == shipped
    ev = np.load(d / "greedy_evaluations.npz", allow_pickle=True)
    assert ev["timesteps"].tolist() == [10000] and ev["actions"].shape[:2] == (1, 30) and ev["rewards"].shape == (1, 30)
    assert "++++++++ Evaluation info (adaptive IQN) ++++++++" in r.stdout and "======== training info ========" in r.stdout
    import torch
    sd = torch.load(d / "network_params.pth", map_location="cpu")
    assert list(sd.keys())[0] == "velocity_encoder.weight" and len(sd) == 14
