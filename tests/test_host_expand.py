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
    """What mnv_pack_obs writes: head, the beam masks, and per group of 32 environments a contiguous run of (x, y) values
    (environment by environment, beam by beam) that starts at dir[g]; the groups' runs sit in the list in any order."""
    E, D = obs.shape
    nb = (D - 4) // 2
    W, G = (nb + 31) // 32, (E + 31) // 32
    pts = obs[:, 4:].reshape(E, nb, 2)
    hit = (pts[..., 0] != 0) | (pts[..., 1] != 0)
    mask = np.zeros((E, W), np.uint32)
    for b in range(nb):
        mask[:, b // 32] |= (hit[:, b].astype(np.uint32) << np.uint32(b % 32))
    dir_ = np.zeros(G, np.uint32)
    vals = np.zeros((int(hit.sum()) + 1, 2), np.float32)
    pos = 0
    for g in rs.permutation(G):
        dir_[g] = pos
        blk = pts[32 * g:32 * g + 32][hit[32 * g:32 * g + 32]]
        vals[pos:pos + len(blk)] = blk
        pos += len(blk)
    return np.ascontiguousarray(obs[:, :4]), mask, dir_, vals


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("E,nb,threads", [(1000, 11, 1), (1000, 11, 3), (4097, 11, 8), (300, 64, 4), (5, 11, 8)])
def test_expand_follows_dense_blocks(E, nb, threads):
    rs = np.random.RandomState(E + threads)
    ex = _hostlib.Expander(E, 4 + 2 * nb, n_threads=threads, cpu_first=-1)
    assert ex.n_threads == min(threads, (E + 31) // 32)
    truth = dense_block(rs, E, nb, 0.3)
    out = truth.copy()
    ex.rescan(ptr(out))                                         # a dense refresh (what reset_host does)
    for step in range(12):
        nxt = dense_block(rs, E, nb, [0.04, 0.5, 0.0, 1.0][step % 4])
        head, mask, dir_, vals = pack(nxt, rs)
        skip = None
        if step % 3 == 1:                                       # auto-reset rows: the "GPU" has written them already
            skip = (rs.rand(E) < 0.1).astype(np.uint8)
            fresh = dense_block(rs, E, nb, 0.2)
            nxt[skip != 0] = fresh[skip != 0]
            out[skip != 0] = fresh[skip != 0]
        ex.expand(ptr(out), ptr(head), None if skip is None else ptr(skip), ptr(mask), ptr(dir_), ptr(vals))
        np.testing.assert_array_equal(out, nxt)
    ex.close()


@pytest.mark.parametrize("E,nb,threads", [(1000, 11, 3), (4097, 11, 8), (300, 64, 4)])
def test_expand_early_then_rescan_skipped(E, nb, threads):
    """The hybrid transport's order: the packet is expanded BEFORE the GPU's rows (auto-reset) have landed in the dense array;
    they land afterwards and rescan_skipped() picks them up -- the following steps must still produce the dense blocks."""
    rs = np.random.RandomState(7 * E + threads)
    ex = _hostlib.Expander(E, 4 + 2 * nb, n_threads=threads, cpu_first=-1)
    out = dense_block(rs, E, nb, 0.3)
    ex.rescan(ptr(out))
    for step in range(10):
        nxt = dense_block(rs, E, nb, [0.04, 0.5, 0.0, 1.0][step % 4])
        head, mask, dir_, vals = pack(nxt, rs)
        skip = (rs.rand(E) < (0.1 if step % 2 else 0.0)).astype(np.uint8)
        fresh = dense_block(rs, E, nb, 0.6)
        nxt[skip != 0] = fresh[skip != 0]
        before = out.copy()
        ex.expand_early(ptr(out), ptr(head), ptr(skip), ptr(mask), ptr(dir_), ptr(vals))
        np.testing.assert_array_equal(out[skip != 0], before[skip != 0])         # flagged rows untouched so far
        out[skip != 0] = fresh[skip != 0]                                        # ... now the "GPU" writes them
        ex.rescan_skipped(ptr(out), ptr(skip))
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
    assert first == -1                                          # no pinning unless MNV_HOST_PIN=1
    monkeypatch.setenv("MNV_HOST_PIN", "1")
    n, first = _hostlib.default_threads_and_first_cpu()
    if cpus == list(range(cpus[0], cpus[0] + n_cpu)) and n <= n_cpu // 2:
        assert first == cpus[0] + max(1, n_cpu // 2)
