"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed over NCCL (gloo in the CPU tests).

The env path shards trivially -- rank g owns the global env indices [g*E_g, (g+1)*E_g) and seeds them by GLOBAL index,
so results do not depend on the world size and `step` needs no collective.  The IQN replicas exchange exactly one
all-reduce of the flat 35 785-float gradient per update (between backward and clip_grad_norm_, agent.py:298-299).
"""
import os

import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def world_size():
    return dist.get_world_size() if is_initialized() else 1


def rank():
    return dist.get_rank() if is_initialized() else 0


def init_from_env(backend=None, device=None):
    """Initialise from torchrun's RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*; no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or is_initialized():
        return rank(), world_size()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        kw["device_id"] = torch.device("cuda", local) if device is None else device
    dist.init_process_group(backend, **kw)
    return rank(), world_size()


def all_reduce_sum_(t):
    """In-place SUM all-reduce of `t` over all ranks; returns the world size (1 and no-op when not distributed)."""
    if not is_initialized() or dist.get_world_size() == 1:
        return 1
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return dist.get_world_size()


def broadcast_(t, src=0):
    if is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src=src)
    return t


def barrier():
    if is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def shard_range(n_total, r=None, world=None):
    """Contiguous 1-D partition of n_total environments: rank r owns [lo, hi)."""
    r = rank() if r is None else r
    world = world_size() if world is None else world
    base, rem = divmod(n_total, world)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def global_env_seeds(base_seed, lo, hi):
    """Seed of global environment index i is base_seed + i (mod 2^32): invariant to how the indices are sharded."""
    return [(int(base_seed) + i) % (1 << 32) for i in range(lo, hi)]
