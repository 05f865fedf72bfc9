"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the reference's own hot-path sources into oracle/_ref/.

The reference (RobustFieldAutonomyLab/Distributional_RL_Navigation @ e77bbbf) is pure Python: there is nothing to
compile, so "building" oracle/_ref means placing the UNMODIFIED files of the hot path where the
GPU box can find them -- /root/reference does not exist there, oracle/_ref/ travels with the gpurun snapshot.  The
directory is git-ignored (never committed: no reference source enters the history) and is only ever imported by
bench.py's CPU-baseline legs (`--impl reference`, `cpu_baseline`) and by tests, through oracle/ref_import.py.

    python oracle/stage_ref.py            # what __graft_entry__.build() runs when /root/reference is present
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("MARINENAV_REF_SRC", "/root/reference")

# the hot path (env step / reset, IQN act / train) + the pretrained weights used as a realistic parameter set
FILES = [
    "marinenav_env/__init__.py",
    "marinenav_env/envs/__init__.py",
    "marinenav_env/envs/marinenav_env.py",
    "marinenav_env/envs/utils/robot.py",
    "thirdparty/IQN/__init__.py",
    "thirdparty/IQN/agent.py",
    "thirdparty/IQN/model.py",
    "thirdparty/IQN/replay_buffer.py",
    "train_IQN_model.py",                    # the caller that must run UNCHANGED on the drop-in packages (tests/test_dropin_driver.py)
    "config/config_IQN.json",
    "pretrained_models/IQN/seed_3/network_params.pth",
    "pretrained_models/IQN/seed_3/constructor_params.json",
    "LICENSE",
]


def stage(verbose=False):
    """Copy FILES from the reference tree; returns DEST, or None when the reference tree is absent (GPU box)."""
    if not os.path.isfile(os.path.join(SRC, FILES[2])):
        return DEST if os.path.isfile(os.path.join(DEST, FILES[2])) else None
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DEST, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.isfile(dst) or os.path.getmtime(src) > os.path.getmtime(dst) or os.path.getsize(src) != os.path.getsize(dst):
            shutil.copy2(src, dst)
            if verbose:
                print("staged", rel)
    with open(os.path.join(DEST, "README"), "w") as f:
        f.write("Unmodified files of RobustFieldAutonomyLab/Distributional_RL_Navigation @ e77bbbf, staged by oracle/stage_ref.py\n"
                "for bench.py's CPU-baseline legs.  Git-ignored; do not edit.\n")
    return DEST


if __name__ == "__main__":
    print(stage(verbose=True))
    sys.exit(0)
