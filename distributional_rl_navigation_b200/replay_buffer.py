"""Replay buffers.

ReplayBuffer      host ring buffer with the interface AND the sampling stream of thirdparty/IQN/replay_buffer.py:6-59
                  (python `random` seeded in the constructor, random.sample over the stored order, n-step folding) --
                  used by the single-env drop-in path so that a run follows the reference's random choices.
DeviceReplayBuffer  device-resident ring buffer for the vectorised trainer, hand-written kernels (csrc/replay.cu): one
                  launch appends a whole env batch (n-step folding included), a Philox draw + gather kernel pair samples
                  (with the without-replacement distribution of random.sample on request); nothing leaves HBM.
"""
import ctypes as C
import random
from collections import deque

import numpy as np
import torch

from . import _lib


class ReplayBuffer:
    def __init__(self, buffer_size, batch_size, device, seed, gamma, n_step=1, state_dim=26):
        self.device = device
        self.capacity = int(buffer_size)
        self.batch_size = batch_size
        self.seed = random.seed(seed)                       # replay_buffer.py:21: seeds the GLOBAL python stream (also used by act)
        self.gamma = gamma
        self.n_step = n_step
        self.n_step_buffer = deque(maxlen=self.n_step)
        self._alloc = 0
        self._states = self._next = self._act = self._rew = self._done = None
        self._head = 0                                      # index of the OLDEST element
        self._size = 0
        self.state_dim = state_dim

    def _grow(self, need):
        new = min(self.capacity, max(1024, 2 * self._alloc, need))
        if new <= self._alloc:
            return

        def grown(a, shape, dtype):
            b = np.zeros((new,) + shape, dtype)
            if a is not None and self._size:
                idx = (self._head + np.arange(self._size)) % self._alloc
                b[:self._size] = a[idx]
            return b
        self._states = grown(self._states, (self.state_dim,), np.float32)
        self._next = grown(self._next, (self.state_dim,), np.float32)
        self._act = grown(self._act, (), np.int64)
        self._rew = grown(self._rew, (), np.float32)
        self._done = grown(self._done, (), np.float32)
        self._head, self._alloc = 0, new

    def _append(self, state, action, reward, next_state, done):
        if self._size == self._alloc and self._alloc < self.capacity:
            self._grow(self._size + 1)
        if self._size < self._alloc:
            pos = (self._head + self._size) % self._alloc
            self._size += 1
        else:                                               # full: drop the oldest (deque(maxlen) semantics)
            pos = self._head
            self._head = (self._head + 1) % self._alloc
        self._states[pos] = state; self._next[pos] = next_state
        self._act[pos] = action; self._rew[pos] = reward; self._done[pos] = float(bool(done))

    def add(self, state, action, reward, next_state, done):
        """replay_buffer.py:26-34"""
        self.n_step_buffer.append((state, action, reward, next_state, done))
        if len(self.n_step_buffer) == self.n_step:
            ret = 0
            for idx in range(self.n_step):                  # calc_multistep_return, replay_buffer.py:36-41
                ret += self.gamma ** idx * self.n_step_buffer[idx][2]
            self._append(self.n_step_buffer[0][0], self.n_step_buffer[0][1], ret, self.n_step_buffer[-1][3], self.n_step_buffer[-1][4])

    def sample(self):
        """replay_buffer.py:45-55: random.sample over the stored order -> (states, actions, rewards, next_states, dones)."""
        picks = random.sample(range(self._size), k=self.batch_size)      # same index stream as random.sample(deque, k)
        idx = (self._head + np.asarray(picks)) % self._alloc
        to = lambda a: torch.from_numpy(a).to(self.device)
        return (to(self._states[idx]), to(self._act[idx].reshape(-1, 1)), to(self._rew[idx].reshape(-1, 1)),
                to(self._next[idx]), to(self._done[idx].reshape(-1, 1)))

    def __len__(self):
        return self._size


class DeviceReplayBuffer:
    """thirdparty/IQN/replay_buffer.py for E parallel environments, resident in HBM (kernels: csrc/replay.cu).

    add_batch  = ReplayBuffer.add for one vector step (per-environment n-step window + fold, ring append): ONE launch.
    sample     = ReplayBuffer.sample: uniform picks (without_replacement=True: the distribution of random.sample) from a
                 counter-based device stream + gather into the (states, actions, rewards, next_states, dones) batch: two
                 launches, nothing leaves the device.  sample(indices=...) gathers caller-provided logical indices
                 (index 0 = the oldest stored transition, like indexing the reference's deque).
    len(buf)   = number of stored transitions."""

    def __init__(self, buffer_size, batch_size, device, seed=0, state_dim=26, gamma=0.99, n_step=1, num_envs=None):
        self.capacity, self.batch_size, self.device = int(buffer_size), int(batch_size), torch.device(device)
        self.gamma, self.n_step, self.state_dim, self.seed = float(gamma), int(n_step), int(state_dim), int(seed)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.states = torch.zeros(self.capacity, state_dim, **f32)
        self.next_states = torch.zeros(self.capacity, state_dim, **f32)
        self.actions = torch.zeros(self.capacity, dtype=torch.int64, device=self.device)
        self.rewards = torch.zeros(self.capacity, **f32)
        self.dones = torch.zeros(self.capacity, **f32)
        self.pos, self.size, self.t, self.calls = 0, 0, 0, 0   # next write slot, stored transitions, vector steps appended, samples drawn
        self._win = None                                         # n-step windows, allocated at the first add_batch (needs E)
        self._out = {}

    @property
    def head(self):
        """Ring slot of the oldest stored transition."""
        return (self.pos - self.size) % self.capacity

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def fill_ctl(self, c):
        """Write the ring state the NEXT add_batch / sample would pass by value into the mnv_vstep_ctl `c` (host struct)."""
        c.rpl_pos, c.rpl_t, c.rpl_head, c.rpl_size, c.rpl_call = self.pos, self.t, self.head, self.size, self.calls

    def advance_append(self, E):
        """Host-side bookkeeping of one add_batch (called by add_batch; by the graph trainer after a replay)."""
        if self.t >= self.n_step - 1:                            # the windows are full: E folded transitions were stored
            self.pos = (self.pos + E) % self.capacity
            self.size = min(self.capacity, self.size + E)
        self.t += 1

    def add_batch(self, states, actions, rewards, next_states, dones, ctl=None, advance=True):
        """Append the transitions of one vector step (device tensors: f32 [E, D], i32 [E], f32 [E], f32 [E, D], u8 [E]).
        ctl (device address of an mnv_vstep_ctl): pos / t come from the control block (rpl_append_ctl, CUDA-graph replays);
        advance=False leaves the host counters alone (a captured launch runs at replay time, not now)."""
        E = states.shape[0]
        if E > self.capacity:
            raise ValueError("batch larger than the buffer")
        if actions.dtype != torch.int32:
            actions = actions.to(torch.int32)
        if dones.dtype != torch.uint8:
            dones = dones.to(torch.uint8)
        for t, dt in ((states, torch.float32), (rewards, torch.float32), (next_states, torch.float32)):
            if t.dtype != dt or not t.is_contiguous() or not t.is_cuda:
                raise _lib.MarinenavError("add_batch: expected contiguous CUDA float32 tensors")
        if self.n_step > 1 and self._win is None:
            self._win = (torch.zeros(self.n_step, E, self.state_dim, dtype=torch.float32, device=self.device),
                         torch.zeros(self.n_step, E, dtype=torch.int32, device=self.device),
                         torch.zeros(self.n_step, E, dtype=torch.float32, device=self.device))
        if self._win is not None and self._win[1].shape[1] != E:
            raise ValueError("add_batch: the n-step windows were sized for %d environments" % self._win[1].shape[1])
        w = self._win or (None, None, None)
        p = _lib.ptr
        with torch.cuda.device(self.device):
            if ctl is not None:
                rc = _lib.load().rpl_append_ctl(p(self.states), p(self.actions), p(self.rewards), p(self.next_states), p(self.dones),
                                                self.capacity, p(states), p(actions.contiguous()), p(rewards), p(next_states),
                                                p(dones.contiguous()), E, self.state_dim, self.n_step, self.gamma,
                                                p(w[0]), p(w[1]), p(w[2]), C.c_void_p(ctl), self._stream())
            else:
                rc = _lib.load().rpl_append(p(self.states), p(self.actions), p(self.rewards), p(self.next_states), p(self.dones),
                                            self.capacity, self.pos, p(states), p(actions.contiguous()), p(rewards), p(next_states),
                                            p(dones.contiguous()), E, self.state_dim, self.n_step, self.gamma, self.t,
                                            p(w[0]), p(w[1]), p(w[2]), self._stream())
        _lib.check(rc, "rpl_append")
        if advance:
            self.advance_append(E)

    def _outputs(self, B):
        o = self._out.get(B)
        if o is None:
            f32 = dict(dtype=torch.float32, device=self.device)
            # two sets, used alternately: the batch of the previous sample() may still be read by a running update
            o = self._out[B] = [0, [dict(idx=torch.zeros(B, dtype=torch.int64, device=self.device),
                                         s=torch.zeros(B, self.state_dim, **f32), a=torch.zeros(B, dtype=torch.int64, device=self.device),
                                         r=torch.zeros(B, **f32), n=torch.zeros(B, self.state_dim, **f32), d=torch.zeros(B, **f32))
                                    for _ in range(2)]]
        o[0] ^= 1
        return o[1][o[0]]

    def sample(self, batch_size=None, without_replacement=False, indices=None, ctl=None, advance=True):
        """-> (states [B, D], actions i64 [B], rewards [B], next_states [B, D], dones [B]); self.last_indices = the picks.
        ctl (device address of an mnv_vstep_ctl): head / size / call come from the control block (rpl_sample_ctl)."""
        B = int(batch_size or self.batch_size)
        if self.size == 0 and ctl is None:
            raise ValueError("sample from an empty buffer")
        o = self._outputs(B)
        p = _lib.ptr
        ring = (p(self.states), p(self.actions), p(self.rewards), p(self.next_states), p(self.dones))
        outs = (p(o["s"]), p(o["a"]), p(o["r"]), p(o["n"]), p(o["d"]))
        with torch.cuda.device(self.device):
            if ctl is not None:
                rc = _lib.load().rpl_sample_ctl(*ring, self.capacity, self.seed, int(bool(without_replacement)), p(o["idx"]), *outs, B,
                                                self.state_dim, C.c_void_p(ctl), self._stream())
                _lib.check(rc, "rpl_sample_ctl")
                if advance:
                    self.calls += 1
            elif indices is not None:
                o["idx"].copy_(torch.as_tensor(indices, dtype=torch.int64).reshape(B))
                rc = _lib.load().rpl_gather(*ring, self.capacity, self.head, self.size, p(o["idx"]), *outs, B, self.state_dim, self._stream())
                _lib.check(rc, "rpl_gather")
            else:
                rc = _lib.load().rpl_sample(*ring, self.capacity, self.head, self.size, self.seed, self.calls,
                                            int(bool(without_replacement)), p(o["idx"]), *outs, B, self.state_dim, self._stream())
                _lib.check(rc, "rpl_sample")
                self.calls += 1
        self.last_indices = o["idx"]
        return o["s"], o["a"], o["r"], o["n"], o["d"]

    def __len__(self):
        return self.size
