"""libmnv_host.so (include/mnv_host.h): the native multi-threaded expander of the compact observation packet, against a
numpy model of the dense observation block (row layout of marinenav_env.py:273-326).  No GPU needed: the packets are
built from dense blocks on the host exactly like mnv_pack_obs builds them on the device (tests/test_dropin_gpu.py checks
that kernel and the whole step_host path)."""
import ctypes as C

import numpy as np
import pytest

from distributional_rl_navigation_b200 import _hostlib, build


@pytest.fixture(scope="module", autouse=True)
def built():
    build.build()


def dense_block(rs, E, nb, p_hit):
    obs = np.zeros((E, 4 + 2 * nb), np.float32)
    obs[:, :4] = rs.randn(E, 4)
    hit = rs.rand(E, nb) < p_hit
    pts = rs.randn(E, nb, 2).astype(np.float32) * 5
    pts[~hit] = 0.0
    obs[:, 4:] = pts.reshape(E, 2 * nb)
    return obs


def pack(obs, rs):
    """What mnv_pack_obs writes: head, and (env << 8 | beam, x bits, y bits) per beam with a return, in any order."""
    E, D = obs.shape
    pts = obs[:, 4:].reshape(E, -1, 2)
    e, b = np.nonzero((pts[..., 0] != 0) | (pts[..., 1] != 0))
    order = rs.permutation(len(e))
    e, b = e[order], b[order]
    hits = np.zeros((len(e), 3), np.uint32)
    hits[:, 0] = (e.astype(np.uint32) << 8) | b.astype(np.uint32)
    hits[:, 1:] = pts[e, b].view(np.uint32)
    return np.ascontiguousarray(obs[:, :4]), hits


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("E,nb,threads", [(1000, 11, 1), (1000, 11, 3), (4097, 11, 8), (300, 64, 4), (5, 11, 8)])
def test_expand_follows_dense_blocks(E, nb, threads):
    rs = np.random.RandomState(E + threads)
    ex = _hostlib.Expander(E, 4 + 2 * nb, n_threads=threads, cpu_first=-1)
    assert ex.n_threads == min(threads, E)
    truth = dense_block(rs, E, nb, 0.3)
    out = truth.copy()
    ex.rescan(ptr(out))                                         # a dense refresh (what reset_host does)
    for step in range(12):
        nxt = dense_block(rs, E, nb, [0.04, 0.5, 0.0, 1.0][step % 4])
        head, hits = pack(nxt, rs)
        skip = None
        if step % 3 == 1:                                       # auto-reset rows: the "GPU" has written them already
            skip = (rs.rand(E) < 0.1).astype(np.uint8)
            fresh = dense_block(rs, E, nb, 0.2)
            nxt[skip != 0] = fresh[skip != 0]
            out[skip != 0] = fresh[skip != 0]
        ex.expand(ptr(out), ptr(head), None if skip is None else ptr(skip), ptr(hits), len(hits))
        np.testing.assert_array_equal(out, nxt)
    ex.close()


def test_default_thread_split_by_local_rank(monkeypatch):
    import os
    n_cpu = len(os.sched_getaffinity(0))
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "2"); monkeypatch.setenv("LOCAL_RANK", "1")
    monkeypatch.delenv("MNV_HOST_THREADS", raising=False)
    n, first = _hostlib.default_threads_and_first_cpu()
    assert 1 <= n <= max(1, min(8, n_cpu // 2))
    cpus = sorted(os.sched_getaffinity(0))
    if cpus == list(range(cpus[0], cpus[0] + n_cpu)) and n <= n_cpu // 2:
        assert first == cpus[0] + max(1, n_cpu // 2)
