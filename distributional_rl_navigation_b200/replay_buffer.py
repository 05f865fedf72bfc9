"""Replay buffers.

ReplayBuffer      host ring buffer with the interface AND the sampling stream of thirdparty/IQN/replay_buffer.py:6-59
                  (python `random` seeded in the constructor, random.sample over the stored order, n-step folding) --
                  used by the single-env drop-in path so that a run follows the reference's random choices.
DeviceReplayBuffer  device-resident ring buffer for the vectorised trainer: whole env batches are appended with one
                  copy per field, uniform sampling with torch.randint on the device; nothing leaves HBM.
"""
import random
from collections import deque

import numpy as np
import torch


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
    def __init__(self, buffer_size, batch_size, device, seed=0, state_dim=26):
        self.capacity, self.batch_size, self.device = int(buffer_size), int(batch_size), torch.device(device)
        f32 = dict(dtype=torch.float32, device=self.device)
        self.states = torch.zeros(self.capacity, state_dim, **f32)
        self.next_states = torch.zeros(self.capacity, state_dim, **f32)
        self.actions = torch.zeros(self.capacity, dtype=torch.int64, device=self.device)
        self.rewards = torch.zeros(self.capacity, **f32)
        self.dones = torch.zeros(self.capacity, **f32)
        self.pos, self.size = 0, 0
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))

    def add_batch(self, states, actions, rewards, next_states, dones):
        """Append E transitions (device tensors); wraps around like a ring."""
        E = states.shape[0]
        if E > self.capacity:
            raise ValueError("batch larger than the buffer")
        first = min(E, self.capacity - self.pos)
        for dst, src in ((self.states, states), (self.next_states, next_states), (self.actions, actions.to(torch.int64)),
                         (self.rewards, rewards), (self.dones, dones.to(torch.float32))):
            dst[self.pos:self.pos + first].copy_(src[:first])
            if first < E:
                dst[:E - first].copy_(src[first:])
        self.pos = (self.pos + E) % self.capacity
        self.size = min(self.capacity, self.size + E)

    def sample(self, batch_size=None):
        B = batch_size or self.batch_size
        idx = torch.randint(0, self.size, (B,), device=self.device, generator=self.gen)
        return (self.states[idx], self.actions[idx], self.rewards[idx], self.next_states[idx], self.dones[idx])

    def __len__(self):
        return self.size
