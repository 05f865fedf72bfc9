"""Thin tensor-level wrappers over the env entry points of libmarinenav_b200 (include/marinenav_b200.h).

PyTorch tensors are the memory carrier only: every function takes CUDA tensors laid out as the header describes, passes
raw device pointers + the current CUDA stream through the C-ABI and returns without synchronising.
"""
import ctypes as C

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, dtype, shape, name):
    if t is None:
        raise _lib.MarinenavError(f"{name} is None")
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
        raise _lib.MarinenavError(f"{name}: expected contiguous CUDA {dtype} {tuple(shape)}, got "
                                  f"{t.device} {t.dtype} {tuple(t.shape)} contiguous={t.is_contiguous()}")


def alloc_env_buffers(E, max_c, max_o, n_beams, device):
    """All per-environment arrays of one env batch (layout: include/marinenav_b200.h)."""
    f64 = dict(dtype=torch.float64, device=device)
    # reward | done | info | compact observation packet (head, hit count, hit list: mnv_pack_obs) share ONE allocation
    # (16-byte aligned segments) so that the host boundary ships everything with one copy
    seg = lambda n: (n + 15) // 16 * 16
    o_done, o_info, n_pack = seg(4 * E), seg(4 * E) + seg(E), seg(4 * E) + 2 * seg(E)
    hit_cap = max(1024, E * n_beams // 8)                      # list capacity: 12.5 % of the beam slots (a rollout fills ~4 %)
    W, G = (n_beams + 31) // 32, (E + 31) // 32
    o_head = n_pack
    o_mask = o_head + 16 * E
    o_dir = o_mask + seg(4 * W * E)
    o_cnt = o_dir + seg(4 * G)
    o_vals = o_cnt + 16
    whole = torch.zeros(o_vals + 8 * hit_cap, dtype=torch.uint8, device=device)
    pack = whole[:n_pack]
    return dict(
        host_packet=whole, packet_offsets=dict(done=o_done, info=o_info, head=o_head, mask=o_mask, dir=o_dir, count=o_cnt, vals=o_vals,
                                               hit_cap=hit_cap),
        packet_head=whole[o_head:o_mask].view(torch.float32).view(E, 4),
        packet_mask=whole[o_mask:o_mask + 4 * W * E].view(torch.int32).view(E, W),
        packet_dir=whole[o_dir:o_dir + 4 * G].view(torch.int32), packet_count=whole[o_cnt:o_vals].view(torch.int32),
        packet_vals=whole[o_vals:].view(torch.float32).view(hit_cap, 2),
        rdi_pack=pack, rdi_offsets=torch.tensor([0, o_done, o_info]),
        state=torch.zeros(4, E, **f64), velocity=torch.zeros(2, E, **f64), goal=torch.zeros(2, E, **f64),
        cores=torch.zeros(3 * max_c, E, **f64), obstacles=torch.zeros(3 * max_o, E, **f64),
        start_pose=torch.zeros(4, E, **f64),
        action=torch.zeros(E, dtype=torch.int32, device=device), episode_step=torch.zeros(E, dtype=torch.int32, device=device),
        obs=torch.zeros(E, 4 + 2 * n_beams, dtype=torch.float32, device=device),
        reward=pack[:4 * E].view(torch.float32), done=pack[o_done:o_done + E], info=pack[o_info:o_info + E],
        n_placed=torch.zeros(2, E, dtype=torch.uint8, device=device),
    )


def step(buf, params, action=None, obs=None, trajectory=None):
    """MarineNavEnv.step for the whole batch (one kernel). Results land in buf['obs'|'reward'|'done'|'info']."""
    E = buf["state"].shape[1]
    max_c, max_o = buf["cores"].shape[0] // 3, buf["obstacles"].shape[0] // 3
    action = buf["action"] if action is None else action
    obs = buf["obs"] if obs is None else obs
    _chk(buf["state"], torch.float64, (4, E), "state"); _chk(buf["velocity"], torch.float64, (2, E), "velocity")
    _chk(buf["goal"], torch.float64, (2, E), "goal")
    if action.is_cuda or not action.is_pinned():      # a PINNED host tensor is fine too: the kernel reads it zero-copy (UVA)
        _chk(action, torch.int32, (E,), "action")
    elif action.dtype != torch.int32 or tuple(action.shape) != (E,) or not action.is_contiguous():
        raise _lib.MarinenavError("action: expected a contiguous int32 tensor of shape (E,)")
    _chk(buf["episode_step"], torch.int32, (E,), "episode_step")
    _chk(obs, torch.float32, (E, 4 + 2 * params.n_beams), "obs")
    if trajectory is not None:
        _chk(trajectory, torch.float64, (params.n_substeps, 2, E), "trajectory")
    rc = _lib.load().mnv_step(_lib.ptr(buf["state"]), _lib.ptr(buf["velocity"]), _lib.ptr(buf["goal"]),
                              _lib.ptr(buf["cores"]), _lib.ptr(buf["obstacles"]), _lib.ptr(action),
                              _lib.ptr(buf["episode_step"]), _lib.ptr(obs), _lib.ptr(buf["reward"]),
                              _lib.ptr(buf["done"]), _lib.ptr(buf["info"]), _lib.ptr(trajectory), E, max_c, max_o,
                              C.byref(params), _stream())
    _lib.check(rc, "mnv_step")


def observe(buf, params, mask=None, velocity_from_state=False, obs=None):
    """MarineNavEnv.get_observation for the (masked) batch."""
    E = buf["state"].shape[1]
    max_c, max_o = buf["cores"].shape[0] // 3, buf["obstacles"].shape[0] // 3
    obs = buf["obs"] if obs is None else obs
    _chk(obs, torch.float32, (E, 4 + 2 * params.n_beams), "obs")
    if mask is not None:
        _chk(mask, torch.uint8, (E,), "mask")
    rc = _lib.load().mnv_observe(_lib.ptr(buf["state"]), _lib.ptr(buf["velocity"]), _lib.ptr(buf["goal"]),
                                 _lib.ptr(buf["cores"]), _lib.ptr(buf["obstacles"]), _lib.ptr(mask), _lib.ptr(obs),
                                 E, max_c, max_o, C.byref(params), int(bool(velocity_from_state)), _stream())
    _lib.check(rc, "mnv_observe")


def seed(rng_key, rng_pos, seeds):
    """RandomState(seed) per environment (MarineNavEnv.seed)."""
    E = seeds.shape[0]
    _chk(rng_key, torch.int32, (E, 624), "rng_key"); _chk(rng_pos, torch.int32, (E,), "rng_pos")
    _chk(seeds, torch.int32, (E,), "seeds")   # bit pattern of u32 seeds
    rc = _lib.load().mnv_seed(_lib.ptr(rng_key), _lib.ptr(rng_pos), _lib.ptr(seeds), E, _stream())
    _lib.check(rc, "mnv_seed")


def reset(buf, rng_key, rng_pos, reset_params, mask=None):
    """MarineNavEnv.reset (map generation + robot draws) for the (masked) batch; follow with observe(velocity_from_state=True)."""
    E = buf["state"].shape[1]
    max_c, max_o = buf["cores"].shape[0] // 3, buf["obstacles"].shape[0] // 3
    if mask is not None:
        _chk(mask, torch.uint8, (E,), "mask")
    rc = _lib.load().mnv_reset(_lib.ptr(rng_key), _lib.ptr(rng_pos), _lib.ptr(mask), _lib.ptr(buf["state"]),
                               _lib.ptr(buf["goal"]), _lib.ptr(buf["cores"]), _lib.ptr(buf["obstacles"]),
                               _lib.ptr(buf["start_pose"]), _lib.ptr(buf["episode_step"]), _lib.ptr(buf["n_placed"]),
                               E, max_c, max_o, C.byref(reset_params), _stream())
    _lib.check(rc, "mnv_reset")


def pack_obs(obs, head, mask, dir_, count, vals):
    """Compact packet of an observation block (mnv_pack_obs): head f32 [E, 4], mask i32 [E, W], dir i32 [ceil(E / 32)],
    count i32 [4], vals f32 [cap, 2]."""
    E, D = obs.shape
    W, G = ((D - 4) // 2 + 31) // 32, (E + 31) // 32
    _chk(obs, torch.float32, (E, D), "obs"); _chk(head, torch.float32, (E, 4), "head"); _chk(mask, torch.int32, (E, W), "mask")
    _chk(dir_, torch.int32, (G,), "dir"); _chk(count, torch.int32, (4,), "count"); _chk(vals, torch.float32, (vals.shape[0], 2), "vals")
    rc = _lib.load().mnv_pack_obs(_lib.ptr(obs), E, D, _lib.ptr(head), _lib.ptr(mask), _lib.ptr(dir_), _lib.ptr(count), _lib.ptr(vals),
                                  vals.shape[0], _stream())
    _lib.check(rc, "mnv_pack_obs")


def scatter_rows_host(mask, rows, host_rows):
    """host_rows[e] <- rows[e] for every e with mask[e] != 0; host_rows is a PINNED CPU tensor (device-mapped under UVA)."""
    E, D = rows.shape
    _chk(mask, torch.uint8, (E,), "mask"); _chk(rows, torch.float32, (E, D), "rows")
    if not (host_rows.is_pinned() and host_rows.dtype == torch.float32 and host_rows.is_contiguous() and tuple(host_rows.shape) == (E, D)):
        raise _lib.MarinenavError(f"host_rows: expected a pinned contiguous float32 CPU tensor {(E, D)}")
    rc = _lib.load().mnv_scatter_rows_host(_lib.ptr(mask), _lib.ptr(rows), _lib.ptr(host_rows), E, D, _stream())
    _lib.check(rc, "mnv_scatter_rows_host")
