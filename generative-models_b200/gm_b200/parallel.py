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


def require_single_process(what):
    """Trainers whose train() has no gradient exchange (InfoGAN's three optimizers, BEGAN's K controller loop)
    refuse to run as silently independent replicas under torchrun."""
    if world_size() > 1:
        raise RuntimeError("%s.train is not data-parallel: run it in one process (the engine-level statistics "
                           "exchange exists, the Trainer loop does not drive it)" % what)


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


class PeerComm:
    """Peer-mapped exchange buffers for the fused all-reduce + Adam kernel
    (include/gm_b200.h: gm_comm_*, gm_gan_apply_allreduce).  Every rank allocates one buffer,
    the 64-byte CUDA IPC handles travel through torch.distributed (all_gather on the
    default group), and each rank maps its peers' buffers: afterwards the gradient exchange
    never leaves the kernel.  `nfloats` = the larger of the two nets' parameter counts."""

    def __init__(self, nfloats, group=None):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        self.h = _lib.ctx()
        self.rank, self.world = rank_of(group), world_size(group)
        self._group = group
        self.c = C.c_void_p()
        _lib.check(self.h, _lib.lib().gm_comm_create(self.h, int(nfloats), C.byref(self.c)))
        mine = (C.c_ubyte * 64)()
        _lib.check(self.h, _lib.lib().gm_comm_handle(self.c, mine))
        backend = dist.get_backend(group) if self.world > 1 else "none"
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        local = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
        if self.world > 1:
            allh = [torch.empty_like(local) for _ in range(self.world)]
            dist.all_gather(allh, local, group=group)
        else:
            allh = [local]
        blob = bytes(torch.cat([t.cpu() for t in allh]).tolist())
        _lib.check(self.h, _lib.lib().gm_comm_open(self.c, self.rank, self.world, C.c_char_p(blob)))
        if self.world > 1:
            dist.barrier(group=group)      # every rank has mapped every buffer before the first kernel uses them

    def close(self, collective=True):
        """Collective by default: a rank may only free its buffer when no peer kernel can still be reading it."""
        if self.c is not None:
            torch.cuda.synchronize()
            if collective and self.world > 1 and dist.is_initialized():
                dist.barrier(group=self._group)
            self._lib.lib().gm_comm_destroy(self.c)
            self.c = None


def make_peer_comm(nfloats, group=None):
    """PeerComm on every rank, or None everywhere if any rank could not map its peers (no CUDA IPC /
    peer access): callers then use sum_gradients (NCCL) + apply.  GM_DP=nccl forces the fallback."""
    if world_size(group) == 1 or os.environ.get("GM_DP", "peer") != "peer":
        return None
    comm, ok = None, 1
    try:
        comm = PeerComm(nfloats, group)
    except Exception as exc:           # noqa: BLE001  (any failure -> collective fallback)
        ok = 0
        print("[gm_b200] peer all-reduce unavailable on rank %d: %s" % (rank_of(group), exc), flush=True)
    backend = dist.get_backend(group)
    flag = torch.tensor([ok], dtype=torch.int32,
                        device=torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        if comm is not None:
            comm.close(collective=False)     # some rank has no communicator: nothing ran on it yet
        return None
    return comm
