"""CPU (no GPU needed): the C-ABI library loads, exports every symbol include/marinenav_b200.h declares, rejects bad
arguments before touching the device, and the product package never imports the oracle."""
import ctypes as C
import os
import re

import pytest

from distributional_rl_navigation_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "marinenav_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:mnv|iqn|rpl)_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    syms = declared_symbols()
    assert "mnv_step" in syms and "mnv_observe" in syms and "mnv_reset" in syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/marinenav_b200.h but not exported"
    assert set(_lib.exported_symbols()) == set(syms)
    assert lib.mnv_version() == 100


def test_struct_layouts_match_defaults(lib):
    p = _lib.default_params()
    assert (p.dt, p.n_substeps, p.n_beams, p.max_episode_steps) == (0.1, 10, 11, 1000)
    assert abs(p.k_drag - 0.2) < 1e-15 and p.height == 50.0 and p.goal_reward == 100.0
    r = _lib.default_reset_params()
    assert (r.num_cores, r.num_obs, r.min_start_goal_dis, r.max_speed) == (8, 5, 25.0, 2.0)
    assert list(r.obs_r_range) == [1.0, 3.0] and list(r.goal) == [45.0, 45.0]


def test_argument_errors_before_any_launch(lib):
    p = _lib.default_params()
    ok = C.c_void_p(4096)
    args = [ok] * 12
    assert lib.mnv_step(*args, 0, 4, 8, C.byref(p), None) == -3                      # E == 0
    assert lib.mnv_step(*args, 16, 9, 8, C.byref(p), None) == -4                     # cores over capacity
    assert lib.mnv_step(*args, 16, 4, 33, C.byref(p), None) == -4                    # obstacles over capacity
    bad = list(args); bad[0] = None
    assert lib.mnv_step(*bad, 16, 4, 8, C.byref(p), None) == -1                      # null
    bad = list(args); bad[0] = C.c_void_p(4100)
    assert lib.mnv_step(*bad, 16, 4, 8, C.byref(p), None) == -2                      # misaligned
    assert b"aligned" in lib.mnv_last_error_string()
    p.n_beams = 500
    assert lib.mnv_step(*args, 16, 4, 8, C.byref(p), None) == -4
    with pytest.raises(_lib.MarinenavError):
        _lib.check(-4, "mnv_step")


def test_kernel_options_roundtrip(lib):
    for key in (b"tma", b"pdl"):
        old = lib.mnv_get_option(key)
        assert old in (0, 1)
        assert lib.mnv_set_option(key, 1 - old) == 0 and lib.mnv_get_option(key) == 1 - old
        assert lib.mnv_set_option(key, old) == 0
    assert lib.mnv_set_option(b"no-such-switch", 1) == -5 and lib.mnv_get_option(b"no-such-switch") == -5
    assert b"unknown key" in lib.mnv_last_error_string()


def test_scatter_rows_host_argument_errors(lib):
    ok = C.c_void_p(4096)
    assert lib.mnv_scatter_rows_host(None, ok, ok, 16, 26, None) == -1
    assert lib.mnv_scatter_rows_host(ok, ok, ok, 0, 26, None) == -3
    assert lib.mnv_scatter_rows_host(ok, C.c_void_p(4100), ok, 16, 26, None) == -2


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "distributional_rl_navigation_b200")
    extra = [os.path.join(ROOT, d) for d in ("marinenav_env", "thirdparty")]
    for base in [pkg] + [d for d in extra if os.path.isdir(d)]:
        for dirpath, _, files in os.walk(base):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle"
                    assert "marinenav_oracle" not in text, f"{f} references the oracle"


def test_bench_algorithmic_bytes_match_survey_table():
    """490 B per env-step at C2 (8 obstacles, 4 cores, 11 beams, fp64 tables), 1490 B at C5 (32 / 4 / 64)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.algorithmic_bytes_per_env_step(4, 8, 11, 8) == 490
    assert bench.algorithmic_bytes_per_env_step(4, 32, 64, 8) == 1490
    assert bench.algorithmic_bytes_per_env_step(4, 8, 11, 4) == 306
