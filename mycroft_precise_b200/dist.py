"""Multi-GPU plumbing: independent audio streams are block-sharded over ranks (one process per
GPU, torch.distributed); the only exchange on the path is the all-reduce of the detection count
(SURVEY 8e).  Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU (tests).
"""
import os


def shard_range(n_streams: int, rank: int, world: int):
    """Contiguous block sharding: stream s lives on rank s // ceil(n/world); state never migrates."""
    per = -(-n_streams // world)
    lo = min(rank * per, n_streams)
    hi = min(lo + per, n_streams)
    return lo, hi


def owner_of(stream: int, n_streams: int, world: int) -> int:
    per = -(-n_streams // world)
    return stream // per


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, local_rank, world).  Single process when WORLD_SIZE is unset or 1."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local, world


class DetectionCounter:
    """Per-rank detection counts -> global sum.  ``local`` is a 1-element (or [k]) integer tensor
    that the device kernels accumulate into; ``all_reduce`` enqueues one SUM all-reduce on the
    current stream (8..24 bytes: latency-bound, overlaps the next tick's MFCC kernel)."""

    def __init__(self, local):
        self.local = local
        self.total = local.clone()

    def all_reduce(self, async_op=False):
        import torch.distributed as dist
        self.total.copy_(self.local)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.total, op=dist.ReduceOp.SUM, async_op=async_op)
        return None
