"""ctypes binding of libmarinenav_b200.so (include/marinenav_b200.h).  There is NO fallback: if the CUDA library is
missing or a call fails, this raises."""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("MNV_LIB") or os.path.join(_PKG, "lib", "libmarinenav_b200.so")   # MNV_LIB: A/B of two builds (lab use)

INFO_STRINGS = ("normal", "too long episode", "collision", "reach goal", "out of boundary")   # marinenav_env.py:243-257
MAX_CORES, MAX_OBSTACLES, MAX_BEAMS = 8, 32, 128


class MnvParams(C.Structure):
    """mnv_params"""
    _fields_ = [
        ("dt", C.c_double), ("n_substeps", C.c_int32),
        ("accel", C.c_double * 3), ("yaw_rate", C.c_double * 3),
        ("k_drag", C.c_double), ("max_speed", C.c_double),
        ("robot_r", C.c_double), ("core_r", C.c_double), ("goal_dis", C.c_double),
        ("timestep_penalty", C.c_double), ("collision_penalty", C.c_double), ("goal_reward", C.c_double),
        ("sonar_range", C.c_double), ("sonar_angle", C.c_double), ("n_beams", C.c_int32),
        ("max_episode_steps", C.c_int32),
        ("set_boundary", C.c_int32), ("width", C.c_double), ("height", C.c_double),
        ("pdl_prefetch", C.c_int32),
    ]


class MnvResetParams(C.Structure):
    """mnv_reset_params"""
    _fields_ = [
        ("width", C.c_double), ("height", C.c_double), ("core_r", C.c_double), ("v_rel_max", C.c_double), ("p", C.c_double),
        ("v_range", C.c_double * 2), ("obs_r_range", C.c_double * 2), ("clear_r", C.c_double),
        ("reset_start_and_goal", C.c_int32), ("start", C.c_double * 2), ("goal", C.c_double * 2),
        ("random_reset_state", C.c_int32), ("init_theta", C.c_double), ("init_speed", C.c_double), ("max_speed", C.c_double),
        ("num_cores", C.c_int32), ("num_obs", C.c_int32), ("min_start_goal_dis", C.c_double),
    ]


class MnvVstepCtl(C.Structure):
    """mnv_vstep_ctl: what changes between two replays of a captured rollout + learn vector step (64 bytes)."""
    _fields_ = [
        ("act_eps", C.c_float), ("adam_step_size", C.c_float), ("adam_inv_sqrt_bc2", C.c_float), ("reserved", C.c_int32),
        ("act_step", C.c_uint64), ("rpl_pos", C.c_int64), ("rpl_t", C.c_int64), ("rpl_head", C.c_int64), ("rpl_size", C.c_int64),
        ("rpl_call", C.c_uint64),
    ]


class MarinenavError(RuntimeError):
    pass


_vp, _i64, _i32 = C.c_void_p, C.c_int64, C.c_int32

_SIGNATURES = {
    "mnv_version": (C.c_int, []),
    "mnv_last_error_string": (C.c_char_p, []),
    "mnv_default_params": (None, [C.POINTER(MnvParams)]),
    "mnv_default_reset_params": (None, [C.POINTER(MnvResetParams)]),
    "mnv_set_option": (C.c_int, [C.c_char_p, _i32]),
    "mnv_get_option": (C.c_int, [C.c_char_p]),
    "mnv_step": (C.c_int, [_vp] * 12 + [_i64, _i32, _i32, C.POINTER(MnvParams), _vp]),
    "mnv_observe": (C.c_int, [_vp] * 7 + [_i64, _i32, _i32, C.POINTER(MnvParams), _i32, _vp]),
    "mnv_seed": (C.c_int, [_vp] * 3 + [_i64, _vp]),
    "mnv_reset": (C.c_int, [_vp] * 10 + [_i64, _i32, _i32, C.POINTER(MnvResetParams), _vp]),
    "mnv_scatter_rows_host": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "mnv_pack_obs": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "rpl_append": (C.c_int, [_vp] * 5 + [_i64, _i64] + [_vp] * 5 + [_i64, _i32, _i32, C.c_double, _i64] + [_vp] * 4),
    "rpl_sample": (C.c_int, [_vp] * 5 + [_i64, _i64, _i64, C.c_uint64, C.c_uint64, _i32] + [_vp] * 6 + [_i64, _i32, _vp]),
    "rpl_gather": (C.c_int, [_vp] * 5 + [_i64, _i64, _i64] + [_vp] * 6 + [_i64, _i32, _vp]),
    "iqn_param_count": (C.c_int, []),
    "iqn_packed_count": (C.c_int, []),
    "iqn_pack": (C.c_int, [_vp, _vp, _vp]),
    "iqn_forward": (C.c_int, [_vp] * 5 + [C.c_float] + [_vp] * 3 + [_i64, _i32, _vp]),
    "iqn_train_scratch_floats": (C.c_int64, [_i64]),
    "iqn_loss_grad": (C.c_int, [_vp] * 11 + [C.c_float] + [_vp] * 3 + [_i64, _vp]),
    "iqn_clip_adam": (C.c_int, [_vp] * 6 + [C.c_float] * 6 + [_i64, _vp, _vp]),
    "iqn_loss_partials": (C.c_int, [_vp] * 11 + [C.c_float] + [_vp] + [_i64, _vp]),
    "iqn_tail_sync_bytes": (C.c_int64, []),
    "iqn_xchg_bytes": (C.c_int64, []),
    "iqn_xchg_handle_bytes": (C.c_int32, []),
    "iqn_xchg_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p]),
    "iqn_xchg_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "iqn_xchg_close": (C.c_int, [_vp]),
    "iqn_xchg_free": (C.c_int, [_vp]),
    "iqn_update_tail": (C.c_int, [_vp] * 6 + [_i64] + [_vp] * 4 + [C.POINTER(C.c_void_p), _i32, _i32] + [C.c_float] * 5 + [_i64, _vp]),
    "iqn_packed_tc_bytes": (C.c_int, []),
    "iqn_pack_tc": (C.c_int, [_vp, _vp, _vp]),
    "iqn_act_scratch_bytes": (C.c_int64, [_i64]),
    "iqn_act_tc": (C.c_int, [_vp] * 5 + [C.c_float] + [_vp] * 4 + [_i64, _i32, _vp]),
    "iqn_act_tc_sample": (C.c_int, [_vp] * 3 + [_i32, _vp, C.c_float, C.c_float, C.c_uint64, C.c_uint64] + [_vp] * 4 + [_i64, _vp]),
    "rpl_append_ctl": (C.c_int, [_vp] * 5 + [_i64] + [_vp] * 5 + [_i64, _i32, _i32, C.c_double] + [_vp] * 5),
    "rpl_sample_ctl": (C.c_int, [_vp] * 5 + [_i64, C.c_uint64, _i32] + [_vp] * 6 + [_i64, _i32, _vp, _vp]),
    "iqn_update_tail_ctl": (C.c_int, [_vp] * 6 + [_i64] + [_vp] * 4 + [C.POINTER(C.c_void_p), _i32, _i32] + [C.c_float] * 4 + [_vp, _vp]),
    "iqn_act_tc_sample_ctl": (C.c_int, [_vp] * 3 + [_i32, _vp, C.c_float, C.c_uint64] + [_vp] * 4 + [_i64, _vp, _vp]),
    "iqn_draw_taus": (C.c_int, [_vp, _i64, C.c_uint64, C.c_uint64, _vp, _vp]),
}

_lib = None


def exported_symbols():
    return list(_SIGNATURES)


def load():
    """Load the CUDA library; raise if it has not been built (python -m distributional_rl_navigation_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(SO_PATH):
            raise MarinenavError(f"{SO_PATH} is missing: build it with `python -m distributional_rl_navigation_b200.build` "
                                 "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)           # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().mnv_last_error_string().decode()
        raise MarinenavError(f"{what} failed (rc={rc}): {msg}")


def set_option(key, value):
    """mnv_set_option: process-wide kernel switches ("tma", "pdl"), see include/marinenav_b200.h."""
    check(load().mnv_set_option(key.encode(), int(value)), f"mnv_set_option({key})")


def get_option(key):
    v = load().mnv_get_option(key.encode())
    if v < 0:
        check(v, f"mnv_get_option({key})")
    return v


def default_params(n_beams=11):
    p = MnvParams()
    load().mnv_default_params(C.byref(p))
    p.n_beams = n_beams
    return p


def default_reset_params():
    p = MnvResetParams()
    load().mnv_default_reset_params(C.byref(p))
    return p


def ptr(t):
    """Device (or host) address of a torch tensor / None."""
    return None if t is None else C.c_void_p(t.data_ptr())
