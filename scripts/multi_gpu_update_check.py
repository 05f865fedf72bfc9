#!/usr/bin/env python
"""Data-parallel IQN update on N GPUs of one node: the fused tail with the one-shot peer-memory all-reduce
(iqn_update_tail) against the three-launch path with NCCL's all-reduce -- same batches, same initial weights.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 scripts/multi_gpu_update_check.py

Checks: replicas bit-identical after every path; the two paths agree (summation order differs: 2e-6 on resolved entries);
timing of both (CUDA events, max over ranks).  Rank 0 prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from distributional_rl_navigation_b200.iqn_agent import IQNAgent

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, n_sets = 1024, 16
g = torch.Generator(device=dev); g.manual_seed(100 + rank)                 # every replica trains on its own batches
st = torch.randn(n_sets, B, 26, device=dev, generator=g) * 3; ns = torch.randn(n_sets, B, 26, device=dev, generator=g) * 3
ac = torch.randint(0, 9, (n_sets, B), device=dev, generator=g); rw = torch.randn(n_sets, B, device=dev, generator=g)
dn = (torch.rand(n_sets, B, device=dev, generator=g) < 0.05).float()
tt = torch.rand(n_sets, B, 8, device=dev, generator=g); tl = torch.rand(n_sets, B, 8, device=dev, generator=g)


def run(fused, n_updates):
    agent = IQNAgent(26, 9, seed=0, device=dev, BATCH_SIZE=B)
    agent.fused_tail = fused
    for i in range(5):
        agent.train_async((st[i % n_sets], ac[i % n_sets], rw[i % n_sets], ns[i % n_sets], dn[i % n_sets]), (tt[i % n_sets], tl[i % n_sets]))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5, 5 + n_updates):
        k = i % n_sets
        agent.train_async((st[k], ac[k], rw[k], ns[k], dn[k]), (tt[k], tl[k]))
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n_updates], device=dev, dtype=torch.float64)
    flat = agent.qnetwork_local.flat.clone()
    identical = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        hi, lo = flat.clone(), flat.clone()
        dist.all_reduce(hi, op=dist.ReduceOp.MAX); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        identical = bool(torch.equal(hi, lo))
    peer = bool(fused and agent._tail is not None and agent._tail.world == world)
    err = None if agent._tail is None else agent._tail.peer_error
    return flat, float(ms.item()), identical, peer, err, float(agent._loss.item())


n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
f_tail, ms_tail, id_tail, peer, err, loss_tail = run(True, n)
f_nccl, ms_nccl, id_nccl, _, _, loss_nccl = run(False, n)
d = (f_tail - f_nccl).abs()
out = {"world": world, "updates": n, "batch_per_gpu": B,
       "fused_tail": {"us_per_update": 1e3 * ms_tail, "replicas_bit_identical": id_tail, "peer_memory_allreduce": peer, "peer_error": err,
                      "last_loss": loss_tail},
       "three_launch_nccl": {"us_per_update": 1e3 * ms_nccl, "replicas_bit_identical": id_nccl, "last_loss": loss_nccl},
       "params_max_abs_diff_between_paths": float(d.max().item()), "params_mean_abs_diff": float(d.mean().item()),
       "lr_times_updates": 1e-4 * (n + 5)}
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
