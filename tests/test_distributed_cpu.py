"""CPU, world_size 2 over gloo: the multi-GPU plumbing (env sharding by global index, the one gradient all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp

from distributional_rl_navigation_b200 import distributed as mdist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w = mdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and mdist.is_initialized()
    # the single data-path collective: SUM all-reduce of the flat gradient, then 1/world scaling in clip_adam
    g = torch.full((35785,), float(rank + 1)) + torch.arange(35785, dtype=torch.float32) * 1e-4
    n = mdist.all_reduce_sum_(g)
    assert n == world
    avg = g / n
    # env sharding: contiguous, complete, seeds keyed by GLOBAL index
    lo, hi = mdist.shard_range(1000003)
    seeds = mdist.global_env_seeds(7, lo, min(hi, lo + 5))
    p = torch.zeros(4) if rank else torch.arange(4.0)
    mdist.broadcast_(p, src=0)
    mdist.barrier()
    torch.save(dict(avg=avg, lo=lo, hi=hi, seeds=seeds, p=p), os.path.join(out_dir, f"r{rank}.pt"))
    torch.distributed.destroy_process_group()


def test_two_rank_allreduce_and_sharding(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"r{i}.pt")) for i in range(2))
    assert torch.equal(r0["avg"], r1["avg"])                            # every rank holds the identical averaged gradient
    want = 1.5 + torch.arange(35785, dtype=torch.float32) * 1e-4
    assert torch.allclose(r0["avg"], want, rtol=1e-6)
    assert (r0["lo"], r0["hi"], r1["lo"], r1["hi"]) == (0, 500002, 500002, 1000003)
    assert r0["seeds"] == [7, 8, 9, 10, 11] and r1["seeds"][0] == 7 + 500002
    assert torch.equal(r1["p"], torch.arange(4.0))


def test_shard_range_is_a_partition_for_any_world():
    for n in (1, 7, 65536, 524288, 131072 + 3):
        for world in (1, 2, 3, 4, 8):
            parts = [mdist.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    assert mdist.world_size() == 1 and mdist.rank() == 0
    t = torch.ones(3)
    assert mdist.all_reduce_sum_(t) == 1 and torch.equal(t, torch.ones(3))


def test_host_replay_buffer_follows_reference_sampling_stream():
    """ReplayBuffer mirrors thirdparty/IQN/replay_buffer.py: deque(maxlen) order + random.sample index stream + n-step."""
    import random
    from collections import deque
    from distributional_rl_navigation_b200.replay_buffer import ReplayBuffer
    rs = np.random.RandomState(0)
    buf = ReplayBuffer(50, 8, "cpu", seed=5, gamma=0.99, n_step=1)
    ref = deque(maxlen=50)
    random.seed(5)
    for t in range(137):
        s, s2 = rs.randn(26), rs.randn(26)
        a, r, d = int(rs.randint(9)), float(rs.randn()), bool(rs.rand() < 0.1)
        buf.add(s, a, r, s2, d); ref.append((s, a, r, s2, d))
    assert len(buf) == 50
    state = random.getstate()
    got = buf.sample()
    random.setstate(state)
    exp = random.sample(ref, k=8)
    np.testing.assert_allclose(got[0].numpy(), np.stack([e[0] for e in exp]).astype(np.float32))
    assert got[1].numpy().ravel().tolist() == [e[1] for e in exp] and got[1].dtype == torch.int64 and got[1].shape == (8, 1)
    np.testing.assert_allclose(got[2].numpy().ravel(), np.float32([e[2] for e in exp]))
    assert got[4].numpy().ravel().tolist() == [float(e[4]) for e in exp]
    # n-step folding (replay_buffer.py:36-41)
    b3 = ReplayBuffer(10, 2, "cpu", seed=0, gamma=0.5, n_step=3)
    for t in range(4):
        b3.add(np.full(26, t), t, 1.0, np.full(26, t + 1), t == 3)
    assert len(b3) == 2 and abs(b3._rew[0] - (1 + 0.5 + 0.25)) < 1e-6 and b3._states[1][0] == 1 and b3._next[1][0] == 4
