"""Device replay buffer (csrc/replay.cu, DeviceReplayBuffer) against the semantics of thirdparty/IQN/replay_buffer.py:
add with n-step folding (:26-41), deque(maxlen) ring (:18), sample = random.sample's distribution (:45-55)."""
import os
from collections import deque

import numpy as np
import pytest
import torch

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    """Plain-Python restatement of Philox4x32-10 (Salmon et al., SC'11) -- the checker of csrc/philox.cuh."""
    x, y, z, w = ctr
    k0, k1 = key
    for _ in range(10):
        p0, p1 = M0 * x, M1 * z
        x, y, z, w = (p1 >> 32) ^ y ^ k0, p1 & 0xffffffff, (p0 >> 32) ^ w ^ k1, p0 & 0xffffffff
        k0, k1 = (k0 + W0) & 0xffffffff, (k1 + W1) & 0xffffffff
    return x, y, z, w


def model_pick(seed, call, pick, rnd, size):
    r = philox4x32_10((pick, rnd, call & 0xffffffff, call >> 32), (seed & 0xffffffff, seed >> 32))
    return (((r[0] << 32) | r[1]) * size) >> 64


def model_sample(seed, call, B, size, without_replacement):
    """The draw kernel's procedure, sequentially: reject a pick iff a lower-numbered pick currently holds its value."""
    picks, rounds = [model_pick(seed, call, j, 0, size) for j in range(B)], [0] * B
    while without_replacement:
        rej = [j for j in range(B) if picks[j] in picks[:j]]
        if not rej:
            break
        for j in rej:
            rounds[j] += 1
            picks[j] = model_pick(seed, call, j, rounds[j], size)
    return picks


def test_philox_known_answers():
    """Random123's published known-answer vectors for philox4x32-10."""
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_model_sampler_is_uniform_over_ordered_distinct_pairs():
    """size 4, B 2: the 12 ordered pairs of distinct indices come out equally often (the distribution of random.sample)."""
    counts = {}
    for call in range(6000):
        p = tuple(model_sample(9, call, 2, 4, True))
        assert p[0] != p[1]
        counts[p] = counts.get(p, 0) + 1
    assert len(counts) == 12 and min(counts.values()) > 400 and max(counts.values()) < 600     # 500 +- 4.5 sigma


class RefModel:
    """replay_buffer.py:6-59 restated for E environments in lockstep: one n-step window per environment, one shared ring."""

    def __init__(self, capacity, gamma, n_step, E):
        self.memory, self.gamma, self.n_step = deque(maxlen=capacity), gamma, n_step
        self.win = [deque(maxlen=n_step) for _ in range(E)]

    def add_batch(self, s, a, r, s2, d):
        for e in range(len(a)):
            w = self.win[e]
            w.append((s[e], a[e], r[e], s2[e], d[e]))                               # :29
            if len(w) == self.n_step:                                               # :30
                ret = 0
                for idx in range(self.n_step):                                      # :36-41
                    ret += self.gamma ** idx * float(w[idx][2])
                self.memory.append((w[0][0], w[0][1], ret, w[-1][3], w[-1][4]))


@pytest.mark.gpu
@pytest.mark.parametrize("n_step,capacity", [(1, 37), (3, 37), (2, 1000)])
def test_append_matches_reference_semantics(n_step, capacity):
    from distributional_rl_navigation_b200.replay_buffer import DeviceReplayBuffer
    E, D, T = 5, 26, 23
    rs = np.random.RandomState(n_step)
    buf = DeviceReplayBuffer(capacity, 8, "cuda:0", seed=1, gamma=0.97, n_step=n_step)
    ref = RefModel(capacity, 0.97, n_step, E)
    dev = lambda x: torch.from_numpy(x).cuda()
    for t in range(T):
        s, s2 = rs.randn(E, D).astype(np.float32), rs.randn(E, D).astype(np.float32)
        a, r, d = rs.randint(0, 9, E).astype(np.int32), rs.randn(E).astype(np.float32), (rs.rand(E) < 0.2).astype(np.uint8)
        buf.add_batch(dev(s), dev(a), dev(r), dev(s2), dev(d))
        ref.add_batch(s, a, r, s2, d)
        assert len(buf) == len(ref.memory)
        if len(buf) == 0:
            continue
        n = len(buf)
        got = [x.cpu().numpy().copy() for x in buf.sample(n, indices=np.arange(n))]     # logical order = deque order
        exp = list(ref.memory)
        np.testing.assert_array_equal(got[0], np.stack([x[0] for x in exp]))
        np.testing.assert_array_equal(got[1], np.array([x[1] for x in exp], np.int64))
        np.testing.assert_allclose(got[2], np.array([x[2] for x in exp], np.float32), rtol=2e-7, atol=0)      # the double sum rounded to float32, like .float() at replay_buffer.py:51
        np.testing.assert_array_equal(got[3], np.stack([x[3] for x in exp]))
        np.testing.assert_array_equal(got[4], np.array([x[4] for x in exp], np.float32))


@pytest.mark.gpu
def test_sample_follows_the_philox_model_and_is_without_replacement():
    from distributional_rl_navigation_b200.replay_buffer import DeviceReplayBuffer
    E, D = 64, 26
    buf = DeviceReplayBuffer(4096, 32, "cuda:0", seed=0xABCDEF0123)
    for t in range(5):                                                           # 320 stored transitions, rewards = logical index
        base = t * E
        z = torch.zeros(E, D, device="cuda")
        buf.add_batch(z + base, torch.zeros(E, dtype=torch.int32, device="cuda"),
                      torch.arange(base, base + E, device="cuda", dtype=torch.float32), z, torch.zeros(E, dtype=torch.uint8, device="cuda"))
    size = len(buf)
    assert size == 320
    for call, (B, wo) in enumerate([(32, True), (300, True), (32, False), (320, True), (1024, False)]):
        s, a, r, s2, d = buf.sample(B, without_replacement=wo)
        idx = buf.last_indices.cpu().numpy()
        assert idx.tolist() == model_sample(buf.seed, call, B, size, wo)          # the device stream IS the Philox model
        assert idx.min() >= 0 and idx.max() < size
        if wo:
            assert len(set(idx.tolist())) == B                                     # random.sample: no repeats (320 of 320 = a permutation)
        np.testing.assert_array_equal(r.cpu().numpy(), idx.astype(np.float32))     # the gathered rows are the picked ones
    a1 = buf.sample(32, without_replacement=True)[2].clone()
    buf.calls -= 1
    a2 = buf.sample(32, without_replacement=True)[2].clone()
    assert torch.equal(a1, a2)                                                     # (seed, call) determines the batch


@pytest.mark.gpu
def test_sample_marginals_are_uniform():
    from distributional_rl_navigation_b200.replay_buffer import DeviceReplayBuffer
    buf = DeviceReplayBuffer(256, 64, "cuda:0", seed=5)
    z = torch.zeros(256, 26, device="cuda")
    buf.add_batch(z, torch.zeros(256, dtype=torch.int32, device="cuda"), torch.zeros(256, device="cuda"), z,
                  torch.zeros(256, dtype=torch.uint8, device="cuda"))
    counts = np.zeros(256)
    for _ in range(400):
        buf.sample(64, without_replacement=True)
        counts += np.bincount(buf.last_indices.cpu().numpy(), minlength=256)
    assert counts.sum() == 400 * 64 and abs(counts - 100).max() < 45               # binomial(400, 1/4): sigma 8.7
