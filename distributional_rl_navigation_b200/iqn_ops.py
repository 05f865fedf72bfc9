"""Tensor-level wrappers over the IQN entry points of libmarinenav_b200 (include/marinenav_b200.h)."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib

N_PARAMS = 35785
N_ENCODER_PARAMS = 4144      # velocity / goal / sensor encoders: the first tensors of the flat vector (iqn_common.cuh oCW)
N_PACKED = 30720
N_ACTIONS = 9
OBS_DIM = 26

# ObsEncoder.state_dict() order and shapes (thirdparty/IQN/model.py:125-136)
PARAM_SPECS = [
    ("velocity_encoder.weight", (16, 2)), ("velocity_encoder.bias", (16,)),
    ("goal_encoder.weight", (16, 2)), ("goal_encoder.bias", (16,)),
    ("sensor_encoder.weight", (176, 22)), ("sensor_encoder.bias", (176,)),
    ("cos_embedding.weight", (208, 64)), ("cos_embedding.bias", (208,)),
    ("hidden_layer.weight", (64, 208)), ("hidden_layer.bias", (64,)),
    ("hidden_layer_2.weight", (64, 64)), ("hidden_layer_2.bias", (64,)),
    ("output_layer.weight", (9, 64)), ("output_layer.bias", (9,)),
]


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32(t, n, name):
    if t is None or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != n:
        raise _lib.MarinenavError(f"{name}: expected contiguous CUDA float32 with {n} elements, got "
                                  f"{None if t is None else (t.device, t.dtype, tuple(t.shape))}")


def pack(params, packed):
    _f32(params, N_PARAMS, "params"); _f32(packed, N_PACKED, "packed")
    _lib.check(_lib.load().iqn_pack(_lib.ptr(params), _lib.ptr(packed), _stream()), "iqn_pack")


def forward(params, packed, obs, taus, cvar=1.0, want_quantiles=True, want_qmean=False, want_greedy=False):
    """ObsEncoder.forward / get_qvals.  obs f32 [B,26], taus f32 [B,n_tau] uniform samples; cvar float or f32 [B]."""
    B, n_tau = taus.shape
    _f32(params, N_PARAMS, "params"); _f32(packed, N_PACKED, "packed")
    _f32(obs, B * OBS_DIM, "obs"); _f32(taus, B * n_tau, "taus")
    dev = obs.device
    cvar_t, cvar_s = (cvar, 1.0) if torch.is_tensor(cvar) else (None, float(cvar))
    if cvar_t is not None:
        _f32(cvar_t, B, "cvar")
    q = torch.empty(B, n_tau, N_ACTIONS, dtype=torch.float32, device=dev) if want_quantiles else None
    qm = torch.empty(B, N_ACTIONS, dtype=torch.float32, device=dev) if want_qmean else None
    gr = torch.empty(B, dtype=torch.int32, device=dev) if want_greedy else None
    rc = _lib.load().iqn_forward(_lib.ptr(params), _lib.ptr(packed), _lib.ptr(obs), _lib.ptr(taus), _lib.ptr(cvar_t),
                                 C.c_float(cvar_s), _lib.ptr(q), _lib.ptr(qm), _lib.ptr(gr), B, n_tau, _stream())
    _lib.check(rc, "iqn_forward")
    return q, qm, gr


def train_scratch_floats(B):
    return int(_lib.load().iqn_train_scratch_floats(B))


def loss_grad(params_l, packed_l, params_t, packed_t, states, actions, rewards, next_states, dones, taus_t, taus_l,
              gamma_n, scratch, loss, grad):
    B = states.shape[0]
    _f32(params_l, N_PARAMS, "params_local"); _f32(params_t, N_PARAMS, "params_target")
    _f32(packed_l, N_PACKED, "packed_local"); _f32(packed_t, N_PACKED, "packed_target")
    _f32(states, B * OBS_DIM, "states"); _f32(next_states, B * OBS_DIM, "next_states")
    _f32(rewards, B, "rewards"); _f32(dones, B, "dones"); _f32(taus_t, B * 8, "taus_target"); _f32(taus_l, B * 8, "taus_local")
    if actions.dtype != torch.int64 or actions.numel() != B or not actions.is_cuda or not actions.is_contiguous():
        raise _lib.MarinenavError("actions: expected contiguous CUDA int64 [B]")
    _f32(grad, N_PARAMS, "grad"); _f32(loss, 1, "loss")
    if scratch.numel() < train_scratch_floats(B):
        raise _lib.MarinenavError("scratch too small")
    rc = _lib.load().iqn_loss_grad(_lib.ptr(params_l), _lib.ptr(packed_l), _lib.ptr(params_t), _lib.ptr(packed_t),
                                   _lib.ptr(states), _lib.ptr(actions), _lib.ptr(rewards), _lib.ptr(next_states),
                                   _lib.ptr(dones), _lib.ptr(taus_t), _lib.ptr(taus_l), C.c_float(gamma_n),
                                   _lib.ptr(scratch), _lib.ptr(loss), _lib.ptr(grad), B, _stream())
    _lib.check(rc, "iqn_loss_grad")


def clip_adam(params, grad, m, v, packed, step, lr=1e-4, max_norm=0.5, grad_scale=1.0, beta1=0.9, beta2=0.999, eps=1e-8,
              grad_norm=None, packed_tc=None):
    """clip_grad_norm_ + Adam.step; `packed` / `packed_tc` (optional) are kept current by the same kernel."""
    for t, n in ((params, "params"), (grad, "grad"), (m, "m"), (v, "v")):
        _f32(t, N_PARAMS, n)
    rc = _lib.load().iqn_clip_adam(_lib.ptr(params), _lib.ptr(grad), _lib.ptr(m), _lib.ptr(v), _lib.ptr(packed), _lib.ptr(packed_tc),
                                   C.c_float(grad_scale), C.c_float(max_norm), C.c_float(lr), C.c_float(beta1),
                                   C.c_float(beta2), C.c_float(eps), int(step), _lib.ptr(grad_norm), _stream())
    _lib.check(rc, "iqn_clip_adam")


def loss_partials(params_l, packed_l, params_t, packed_t, states, actions, rewards, next_states, dones, taus_t, taus_l, gamma_n, scratch):
    """iqn_loss_grad without the reduction: per-tile partial gradients / losses stay in `scratch` (consumed by UpdateTail.step)."""
    B = states.shape[0]
    _f32(params_l, N_PARAMS, "params_local"); _f32(params_t, N_PARAMS, "params_target")
    _f32(packed_l, N_PACKED, "packed_local"); _f32(packed_t, N_PACKED, "packed_target")
    _f32(states, B * OBS_DIM, "states"); _f32(next_states, B * OBS_DIM, "next_states")
    _f32(rewards, B, "rewards"); _f32(dones, B, "dones"); _f32(taus_t, B * 8, "taus_target"); _f32(taus_l, B * 8, "taus_local")
    if actions.dtype != torch.int64 or actions.numel() != B or not actions.is_cuda or not actions.is_contiguous():
        raise _lib.MarinenavError("actions: expected contiguous CUDA int64 [B]")
    if scratch.numel() < train_scratch_floats(B):
        raise _lib.MarinenavError("scratch too small")
    rc = _lib.load().iqn_loss_partials(_lib.ptr(params_l), _lib.ptr(packed_l), _lib.ptr(params_t), _lib.ptr(packed_t),
                                       _lib.ptr(states), _lib.ptr(actions), _lib.ptr(rewards), _lib.ptr(next_states),
                                       _lib.ptr(dones), _lib.ptr(taus_t), _lib.ptr(taus_l), C.c_float(gamma_n),
                                       _lib.ptr(scratch), B, _stream())
    _lib.check(rc, "iqn_loss_partials")


class UpdateTail:
    """Host-side state of iqn_update_tail: the grid-barrier words and, for data-parallel replicas, the peer-mapped exchange
    buffers of the one-shot all-reduce (CUDA IPC handles exchanged once through torch.distributed).

    peer_exchange: None (auto: on when torch.distributed is initialised with more than one rank), True (required) or False
    (single-GPU tail; the caller all-reduces itself).  If the buffers cannot be mapped (no peer access between the GPUs) and
    peer_exchange is None, `self.world` stays 1 and `self.peer_error` says why -- the caller then keeps the NCCL path."""

    def __init__(self, device, peer_exchange=None):
        import torch.distributed as dist
        L = _lib.load()
        self.device = torch.device(device)
        self.sync = torch.zeros(int(L.iqn_tail_sync_bytes()), dtype=torch.uint8, device=self.device)
        self.world, self.rank, self.peer_error = 1, 0, None
        self._own, self._opened, self._peer_array = None, [], None
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if peer_exchange is False or not distributed:
            if peer_exchange is True and not distributed:
                raise _lib.MarinenavError("UpdateTail(peer_exchange=True) needs an initialised torch.distributed group of > 1 ranks")
            return
        world, rank = dist.get_world_size(), dist.get_rank()
        nh = int(L.iqn_xchg_handle_bytes())
        own, handle = C.c_void_p(), C.create_string_buffer(nh)
        with torch.cuda.device(self.device):
            rc = L.iqn_xchg_alloc(C.byref(own), handle)
            err = None if rc == 0 else L.mnv_last_error_string().decode()
            mine = torch.tensor(list(handle.raw) + [0 if rc == 0 else 1], dtype=torch.uint8, device=self.device)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            ptrs = (C.c_void_p * world)()
            ok = all(int(h[-1]) == 0 for h in allh) and world <= 8
            if ok:
                for r in range(world):
                    if r == rank:
                        ptrs[r] = own.value
                        continue
                    q = C.c_void_p()
                    rc = L.iqn_xchg_open(bytes(allh[r][:nh].cpu().tolist()), C.byref(q))
                    if rc != 0:
                        ok, err = False, L.mnv_last_error_string().decode()
                        break
                    ptrs[r] = q.value
                    self._opened.append(q)
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)                 # all ranks or none
            if int(flag.item()) == 1:
                self._own, self._peer_array, self.world, self.rank = own, ptrs, world, rank
            else:
                for q in self._opened:
                    L.iqn_xchg_close(q)
                self._opened = []
                if own.value:
                    L.iqn_xchg_free(own)
                self.peer_error = err or "a peer rank could not map the exchange buffers"
                if peer_exchange is True:
                    raise _lib.MarinenavError(f"UpdateTail: peer exchange unavailable: {self.peer_error}")

    def step(self, params, m, v, packed, packed_tc, scratch, B, step, loss=None, grad=None, grad_norm=None, lr=1e-4, max_norm=0.5,
             beta1=0.9, beta2=0.999, eps=1e-8, ctl=None):
        """ctl (device address of an mnv_vstep_ctl): Adam's bias corrections come from the control block (iqn_update_tail_ctl,
        CUDA-graph replays); `step` and `lr` are then ignored -- see adam_ctl_fields()."""
        for t, n in ((params, "params"), (m, "m"), (v, "v")):
            _f32(t, N_PARAMS, n)
        if grad is not None:
            _f32(grad, N_PARAMS, "grad")
        if ctl is not None:
            rc = _lib.load().iqn_update_tail_ctl(_lib.ptr(params), _lib.ptr(m), _lib.ptr(v), _lib.ptr(packed), _lib.ptr(packed_tc),
                                                 _lib.ptr(scratch), int(B), _lib.ptr(loss), _lib.ptr(grad), _lib.ptr(grad_norm),
                                                 _lib.ptr(self.sync), self._peer_array, self.rank, self.world,
                                                 C.c_float(max_norm), C.c_float(beta1), C.c_float(beta2), C.c_float(eps),
                                                 C.c_void_p(ctl), _stream())
            _lib.check(rc, "iqn_update_tail_ctl")
            return
        rc = _lib.load().iqn_update_tail(_lib.ptr(params), _lib.ptr(m), _lib.ptr(v), _lib.ptr(packed), _lib.ptr(packed_tc),
                                         _lib.ptr(scratch), int(B), _lib.ptr(loss), _lib.ptr(grad), _lib.ptr(grad_norm),
                                         _lib.ptr(self.sync), self._peer_array, self.rank, self.world,
                                         C.c_float(max_norm), C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps),
                                         int(step), _stream())
        _lib.check(rc, "iqn_update_tail")

    def error(self):
        """0, or the sticky error word of the kernel: 1 = a peer's slice never arrived, 2 = the grid barrier timed out (both after
        seconds of spinning; the update that hit it is garbage).  Host-synchronising read: call it outside hot loops."""
        return int(self.sync[16:24].view(torch.int64).item())

    def close(self):
        L = _lib.load()
        if self._own is not None:
            torch.cuda.synchronize(self.device)
            for q in self._opened:
                L.iqn_xchg_close(q)
            L.iqn_xchg_free(self._own)
            self._own, self._opened, self._peer_array, self.world = None, [], None, 1


def packed_tc_bytes():
    return int(_lib.load().iqn_packed_tc_bytes())


def pack_tc(params, packed_tc):
    """fp32 parameters -> bf16 UMMA weight tiles for act_tc (call after any parameter change)."""
    _f32(params, N_PARAMS, "params")
    if packed_tc.dtype != torch.uint8 or packed_tc.numel() != packed_tc_bytes() or not packed_tc.is_cuda:
        raise _lib.MarinenavError("packed_tc: expected CUDA uint8 buffer of iqn_packed_tc_bytes() bytes")
    _lib.check(_lib.load().iqn_pack_tc(_lib.ptr(params), _lib.ptr(packed_tc), _stream()), "iqn_pack_tc")


_act_scratch = {}


def _scratch_for(B, dev):
    """bf16 encoder-feature scratch of the acting kernels, cached per (device, size)."""
    need = int(_lib.load().iqn_act_scratch_bytes(B))
    key = (dev.index if dev.index is not None else torch.cuda.current_device())
    buf = _act_scratch.get(key)
    if buf is None or buf.numel() < need:
        buf = _act_scratch[key] = torch.empty(need, dtype=torch.uint8, device=dev)
    return buf


def act_tc(params, packed_tc, obs, taus, cvar=1.0, want_qmean=False, want_greedy=True, debug=None):
    """get_qvals + argmax on the tensor cores (bf16 operands): obs f32 [B,26], taus f32 [B,32]."""
    B, n_tau = taus.shape
    _f32(params, N_PARAMS, "params"); _f32(obs, B * OBS_DIM, "obs"); _f32(taus, B * n_tau, "taus")
    dev = obs.device
    cvar_t, cvar_s = (cvar, 1.0) if torch.is_tensor(cvar) else (None, float(cvar))
    if cvar_t is not None:
        _f32(cvar_t, B, "cvar")
    qm = torch.empty(B, N_ACTIONS, dtype=torch.float32, device=dev) if want_qmean else None
    gr = torch.empty(B, dtype=torch.int32, device=dev) if want_greedy else None
    rc = _lib.load().iqn_act_tc(_lib.ptr(params), _lib.ptr(packed_tc), _lib.ptr(obs), _lib.ptr(taus), _lib.ptr(cvar_t),
                                C.c_float(cvar_s), _lib.ptr(qm), _lib.ptr(gr), _lib.ptr(debug), _lib.ptr(_scratch_for(B, dev)),
                                B, n_tau, _stream())
    _lib.check(rc, "iqn_act_tc")
    return qm, gr


def adam_ctl_fields(lr, beta1, beta2, step):
    """(adam_step_size, adam_inv_sqrt_bc2) of mnv_vstep_ctl for optimizer step `step` (>= 1): the arithmetic of
    iqn_update_tail's by-value path (csrc/iqn_tail.cu) -- lr / beta1 / beta2 arrive there as floats, the bias corrections are
    formed in double and rounded to float last (the ctypes float fields do that rounding)."""
    lr, beta1, beta2 = (float(np.float32(x)) for x in (lr, beta1, beta2))
    bc1, bc2 = 1.0 - math.pow(beta1, float(int(step))), 1.0 - math.pow(beta2, float(int(step)))
    return lr / bc1, 1.0 / math.sqrt(bc2)


def draw_taus(out, seed, call=0, ctl=None):
    """The quantile samples of one update on the device (iqn_draw_taus): out f32 [2, B, 8] <- uniform [0, 1), reproducible per
    (seed, call); out[0] = the target network's taus (drawn first, Q9), out[1] = the local network's."""
    _f32(out, out.numel(), "taus")
    rc = _lib.load().iqn_draw_taus(_lib.ptr(out), out.numel(), int(seed) & (2 ** 64 - 1), int(call) & (2 ** 64 - 1),
                                   None if ctl is None else C.c_void_p(ctl), _stream())
    _lib.check(rc, "iqn_draw_taus")
    return out


def act_tc_sample(params, packed_tc, obs, eps, seed, step, cvar=1.0, adaptive=False, action=None, want_greedy=False, want_qmean=False,
                  cvar_out=None, ctl=None):
    """IQNAgent.act / act_adaptive for an env batch with taus and the epsilon-greedy draw from the device Philox stream
    (seed, step): -> (action i32 [B], greedy i32 [B] | None, qmean f32 [B, 9] | None).  `params`: the flat parameter vector,
    or just its first N_ENCODER_PARAMS floats (the observation encoders -- all the pre-pass reads)."""
    B = obs.shape[0]
    if not (params.is_cuda and params.dtype == torch.float32 and params.is_contiguous() and params.numel() >= N_ENCODER_PARAMS):
        raise _lib.MarinenavError("params: expected a contiguous CUDA float32 vector holding at least the encoder parameters")
    _f32(obs, B * OBS_DIM, "obs")
    dev = obs.device
    action = torch.empty(B, dtype=torch.int32, device=dev) if action is None else action
    gr = torch.empty(B, dtype=torch.int32, device=dev) if want_greedy else None
    qm = torch.empty(B, N_ACTIONS, dtype=torch.float32, device=dev) if want_qmean else None
    if adaptive and cvar_out is None:
        cvar_out = torch.empty(B, dtype=torch.float32, device=dev)
    if ctl is not None:                # eps / step from the device control block (CUDA-graph replays)
        rc = _lib.load().iqn_act_tc_sample_ctl(_lib.ptr(params), _lib.ptr(packed_tc), _lib.ptr(obs), int(bool(adaptive)), _lib.ptr(cvar_out),
                                               C.c_float(float(cvar)), int(seed) & (2 ** 64 - 1), _lib.ptr(action), _lib.ptr(gr), _lib.ptr(qm),
                                               _lib.ptr(_scratch_for(B, dev)), B, C.c_void_p(ctl), _stream())
        _lib.check(rc, "iqn_act_tc_sample_ctl")
        return action, gr, qm
    rc = _lib.load().iqn_act_tc_sample(_lib.ptr(params), _lib.ptr(packed_tc), _lib.ptr(obs), int(bool(adaptive)), _lib.ptr(cvar_out),
                                       C.c_float(float(cvar)), C.c_float(float(eps)), int(seed) & (2 ** 64 - 1), int(step) & (2 ** 64 - 1),
                                       _lib.ptr(action), _lib.ptr(gr), _lib.ptr(qm), _lib.ptr(_scratch_for(B, dev)), B, _stream())
    _lib.check(rc, "iqn_act_tc_sample")
    return action, gr, qm
