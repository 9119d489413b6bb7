"""N>1 host logic on CPU (gloo, world_size 2): sharding by samples + pre-scaling by
1/(global batch) + SUM all-reduce of the flat gradients reproduces the single-process
global-batch gradient.  Per-rank gradients come from the numpy oracle (no GPU here)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from inputs import GAN_SHAPES, gm_init_weights, gm_images, params_dict
from oracle import ref_math as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "generative-models_b200"))
    from gm_b200 import parallel as P
    r, w, _ = P.init_from_env("gloo")
    assert (r, w) == (rank, world)
    Bg = 64
    P64 = params_dict(gm_init_weights(GAN_SHAPES, 1234), np.float64)
    x = gm_images(Bg).astype(np.float64)
    z = np.random.default_rng(3).standard_normal((Bg, 20))
    lo, hi = P.shard_slice(Bg, rank, world)
    # local mean-loss gradient * (local/global) == per-sample grads scaled by 1/global
    L, g, _ = R.gan_d_step(P64, "ns", x[lo:hi], z[lo:hi])
    scale = (hi - lo) * P.inv_global_batch(hi - lo, world)
    flat = torch.from_numpy(np.concatenate([g[k].ravel() for k in sorted(g)]) * scale)
    P.sum_gradients(flat)
    loss = P.mean_scalar(torch.tensor([L]))
    if rank == 0:
        Lf, gf, _ = R.gan_d_step(P64, "ns", x, z)
        ref = np.concatenate([gf[k].ravel() for k in sorted(gf)])
        out.put((float(np.linalg.norm(flat.numpy() - ref) / np.linalg.norm(ref)), float(abs(loss.item() - Lf))))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sum_equals_global_batch_gradient():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    err, lerr = out.get()
    assert err < 1e-12 and lerr < 1e-12, (err, lerr)


def test_helpers():
    import gm_b200.parallel as P
    assert P.inv_global_batch(64, 8) == 1.0 / 512
    assert P.rank_seed(7, 0) != P.rank_seed(7, 1)
    assert P.shard_slice(128, 1, 2) == (64, 128)
    t = torch.ones(3)
    assert P.sum_gradients(t) is t


def _stats_worker(rank, world, port, out):
    """The device protocol of gm_gan_attach_comm restated on the host: partial sums of the batch
    statistics are exchanged (SUM) between the loss passes, per-row upstream gradients use the
    global statistics and 1/(global batch), and the flat gradients are summed."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "generative-models_b200"))
    from gm_b200 import parallel as P
    P.init_from_env("gloo")
    Bg = 64
    P64 = params_dict(gm_init_weights(GAN_SHAPES, 1234), np.float64)
    x = gm_images(Bg).astype(np.float64)
    z = np.random.default_rng(3).standard_normal((Bg, 20))
    lo, hi = P.shard_slice(Bg, rank, world)
    res = {}
    for variant in ("ra", "fisher"):
        gf = R.g_forward(P64, z[lo:hi])
        fx, fg = R.d_forward(P64, x[lo:hi]), R.d_forward(P64, gf["out"])
        dx, dg = fx["d"], fg["d"]

        def allsum(*vals):
            t = torch.tensor([float(v) for v in vals], dtype=torch.float64)
            dist.all_reduce(t)
            return t.tolist()
        if variant == "ra":
            (sdg,) = allsum(dg.sum())                                   # pass 0 -> exchange
            q = R.sigmoid(dx - sdg / Bg)
            r = R.sigmoid(1 - dg)
            gq = q * (1 - q) / (q + R.EPS)
            (sgq,) = allsum(gq.sum())                                   # pass 1 -> exchange
            ddx = -gq / (2 * Bg)
            ddg = (sgq / Bg + r * (1 - r) / (r + R.EPS)) / (2 * Bg)
            st_full = None
        else:
            lam, rho = 0.3, 1e-2
            s2x, s2g = allsum((dx ** 2).sum(), (dg ** 2).sum())         # pass 0 -> exchange
            om = 1 - (0.5 * s2x / Bg + 0.5 * s2g / Bg)
            c = lam - rho * om
            ddx, ddg = -(1 - c * dx) / Bg, (1 + c * dg) / Bg
            st_full = dict(LAMBDA=lam, RHO=rho)
        g1, _ = R.d_backward(P64, fx, R.d_out_grad(fx, ddx, "sigmoid"))
        g2, _ = R.d_backward(P64, fg, R.d_out_grad(fg, ddg, "sigmoid"))
        flat = torch.from_numpy(np.concatenate([(g1[k] + g2[k]).ravel() for k in sorted(g1)]))
        P.sum_gradients(flat)
        if rank == 0:
            _, gfull, _ = R.gan_d_step(P64, variant, x, z, st=st_full)
            ref = np.concatenate([gfull[k].ravel() for k in sorted(gfull)])
            res[variant] = float(np.linalg.norm(flat.numpy() - ref) / np.linalg.norm(ref))
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_exchanged_batch_statistics_reproduce_the_global_batch():
    """RaNS mean(DG) / sum q(1-q)/(q+eps) and the Fisher moments (src/ra_gan.py:204, src/fisher_gan.py:214-218)
    summed over ranks between the loss passes: 2 ranks x 32 rows give the gradient of 1 process x 64 rows."""
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_stats_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = out.get()
    assert res["ra"] < 1e-12 and res["fisher"] < 1e-12, res
