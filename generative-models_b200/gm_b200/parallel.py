"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink on
the GPU box, gloo in CPU tests).  The train step shards by samples: each
rank runs the fused D / G gradient kernels on its own batch with upstream gradients
pre-scaled by 1/(global batch), so the only exchange is a SUM all-reduce of the flat
D gradient (314 401 floats) after the D backward and of the flat G gradient (322 784
floats) after the G backward — G and D gradients only, nothing else crosses NVLink."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's RANK/WORLD_SIZE/MASTER_* env.
    Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def world_size(group=None):
    """Ranks in the (default) process group; 1 when torch.distributed is not initialised."""
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank_of(group=None):
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def inv_global_batch(local_batch, world):
    """Scale for per-sample upstream gradients so that SUM over ranks == mean over the
    global batch (the reference's losses are batch means, src/ns_gan.py:191-192)."""
    return 1.0 / float(local_batch * world)


def rank_seed(base_seed, rank):
    """Distinct Philox key per rank: every rank draws its own noise."""
    return int(base_seed) * 1000003 + int(rank)


def sum_gradients(flat_grad, group=None):
    """In-place SUM all-reduce of one net's flat gradient buffer; no-op on 1 rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def mean_scalar(t, group=None):
    """Average a logging scalar (loss) across ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t /= dist.get_world_size(group)
    return t


def shard_slice(n, rank, world):
    """Rows [lo, hi) of a global batch of n that rank owns (contiguous, equal shards)."""
    per = n // world
    return rank * per, (rank + 1) * per
