"""Correctness of the fused all-reduce + Adam kernel (gm_gan_apply_allreduce) against
torch.distributed.all_reduce + gm_gan_apply.  Launch one process per rank:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_allreduce_check.py
With GM_PEER_SAME_GPU=1 every rank uses cuda:0 (CUDA IPC works between processes on one GPU, the
two kernels then take turns through context time-slicing) and the rendezvous runs over gloo."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "generative-models_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import gm_b200  # noqa: E402
from gm_b200 import parallel as par  # noqa: E402
from inputs import GAN_SHAPES, gm_init_weights  # noqa: E402

same_gpu = os.environ.get("GM_PEER_SAME_GPU") == "1"
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = 0 if same_gpu else int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("gloo" if same_gpu else "nccl")
B, STEPS = 256, 3


def engine():
    eng = gm_b200.GanEngine(784, 400, 20, max_batch=B, variant="ns")
    W = gm_init_weights(GAN_SHAPES, 1234)
    eng.load(0, [W["G.linear"][0], W["G.linear"][1], W["G.generate"][0], W["G.generate"][1]])
    eng.load(1, [W["D.linear"][0], W["D.linear"][1], W["D.discriminate"][0], W["D.discriminate"][1]])
    return eng


g = torch.Generator().manual_seed(100 + rank)          # every rank its own batch and noise
xs = [(torch.rand(B, 784, generator=g) < 0.13).float().cuda() for _ in range(STEPS)]
zs = [torch.randn(B, 20, generator=g).cuda() for _ in range(2 * STEPS)]
hp = gm_b200.AdamHP.make(2e-4)
inv = par.inv_global_batch(B, world)


def reduce_ref(t):
    if same_gpu:
        c = t.cpu()
        dist.all_reduce(c)
        t.copy_(c)
    else:
        dist.all_reduce(t)


ref = engine()
for s in range(STEPS):
    ref.d_grad(xs[s], noise=zs[2 * s], inv_global_batch=inv, step=s)
    reduce_ref(ref.grads[1])
    ref.apply(1, hp)
    ref.g_grad(B, noise=zs[2 * s + 1], inv_global_batch=inv, step=s)
    reduce_ref(ref.grads[0])
    ref.apply(0, hp)
torch.cuda.synchronize()

eng = engine()
eng.set_lazy_grads(True)
comm = par.PeerComm(max(eng.n))
for s in range(STEPS):
    eng.d_grad(xs[s], noise=zs[2 * s], inv_global_batch=inv, step=s)
    eng.apply_allreduce(1, hp, comm)
    eng.g_grad(B, noise=zs[2 * s + 1], inv_global_batch=inv, step=s)
    eng.apply_allreduce(0, hp, comm)
torch.cuda.synchronize()

worst = 0.0
for net in (0, 1):
    a, b = eng.params[net], ref.params[net]
    worst = max(worst, float((a - b).abs().max() / b.abs().max()))
    ga, gb = eng.grads[net], ref.grads[net]
    worst = max(worst, float((ga - gb).abs().max() / gb.abs().max()))
    # replicas stay bitwise identical: every rank sums the chunks in rank order
    mine = eng.params[net].cpu() if same_gpu else eng.params[net].clone()
    allp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    assert all(torch.equal(allp[0], t) for t in allp), "replicas diverged"
assert worst < 1e-5, worst
# the exchange in two halves (gm_gan_exchange_begin ... gm_gan_apply_allreduce) with independent work in between - the G step's
# generator forward under the D exchange, the next step's image staging under the G exchange - gives bitwise the same replicas
eng2 = engine()
eng2.set_lazy_grads(True)
eng2.d_stage(xs[0], step=0)
for s in range(STEPS):
    eng2.d_grad(xs[s], noise=zs[2 * s], inv_global_batch=inv, step=s)
    eng2.exchange_begin(1, comm)
    eng2.g_forward_stage(B, noise=zs[2 * s + 1], step=s)
    eng2.apply_allreduce(1, hp, comm)
    eng2.g_grad_staged(B, inv_global_batch=inv)
    eng2.exchange_begin(0, comm)
    if s + 1 < STEPS:
        eng2.d_stage(xs[s + 1], step=s + 1)
    eng2.apply_allreduce(0, hp, comm)
torch.cuda.synchronize()
for net in (0, 1):
    assert torch.equal(eng2.params[net], eng.params[net]), "two-phase exchange differs from the fused one"
comm.close()
if rank == 0:
    print("PEER_ALLREDUCE_OK world=%d worst_rel=%.3g" % (world, worst))
    print("SPLIT_EXCHANGE_OK")


# ---- global-batch equivalence: `world` ranks x Bl rows == one engine with world*Bl rows ----------
def global_batch_check(variant):
    comm2 = par.PeerComm(400000)
    Bl = 128
    Bg = Bl * world
    gg = torch.Generator().manual_seed(4242)              # same global tensors on every rank
    X = (torch.rand(Bg, 784, generator=gg) < 0.13).float().cuda()
    Z1, Z2 = torch.randn(Bg, 20, generator=gg).cuda(), torch.randn(Bg, 20, generator=gg).cuda()
    delta, u = torch.rand(Bg, generator=gg).cuda(), torch.rand(Bg, 784, generator=gg).cuda()
    lo, hi = rank * Bl, (rank + 1) * Bl

    def mk(batch):
        e = gm_b200.GanEngine(784, 400, 20, max_batch=batch, variant=variant, d_out_act="relu" if variant == "wgp" else "sigmoid")
        W = gm_init_weights(GAN_SHAPES, 1234)
        e.load(0, [W["G.linear"][0], W["G.linear"][1], W["G.generate"][0], W["G.generate"][1]])
        e.load(1, [W["D.linear"][0], W["D.linear"][1], W["D.discriminate"][0], W["D.discriminate"][1]])
        if variant == "fisher":
            e.fisher_state(0.3, 1e-2)
        return e

    def aux(a, b):
        if variant == "dra":
            return torch.cat([delta[a:b].reshape(-1), u[a:b].reshape(-1)]).contiguous()
        if variant == "wgp":
            return delta[a:b].contiguous()
        return None

    one = mk(Bg)                                           # the single-process run over the global batch
    one.d_grad(X, noise=Z1, aux=aux(0, Bg), inv_global_batch=1.0 / Bg)
    gD_one = one.grads[1].clone()
    loss_one = float(one.loss_buf[0])
    dp = mk(Bl)
    dp.attach_comm(comm2)
    dp.set_lazy_grads(True)
    dp.d_grad(X[lo:hi].contiguous(), noise=Z1[lo:hi].contiguous(), aux=aux(lo, hi), inv_global_batch=1.0 / Bg)
    hp0 = gm_b200.AdamHP.make(0.0)                         # lr 0: only the summed gradient is of interest
    dp.apply_allreduce(1, hp0, comm2)
    torch.cuda.synchronize()
    rel = float((dp.grads[1] - gD_one).norm() / gD_one.norm())
    # per-rank statistics (no communicator) must NOT reproduce it for the batch-statistic variants
    assert rel < 2e-4, (variant, rel)
    if variant == "fisher":
        lam_dp, lam_one = dp.fisher_state()[0], one.fisher_state()[0]
        assert abs(lam_dp - lam_one) < 1e-6 * max(1.0, abs(lam_one)), (lam_dp, lam_one)
    comm2.close()
    return rel


rels = {v: global_batch_check(v) for v in os.environ.get("GM_CHECK_VARIANTS", "ns,ra,fisher,dra").split(",") if v}
if rank == 0:
    print("GLOBAL_BATCH_OK " + " ".join("%s=%.2g" % kv for kv in rels.items()))


# ---- Trainer.train under data parallelism: the drop-in class itself exchanges gradients ----------
def trainer_check():
    sys.path.insert(0, os.path.join(ROOT, "generative-models_b200"))
    import ns_gan
    torch.manual_seed(11)                                   # same initial weights on every rank
    model = ns_gan.NSGAN(784, 400, 20)
    g2 = torch.Generator().manual_seed(5)
    imgs = (torch.rand(512, 1, 28, 28, generator=g2) < 0.13).float()
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(imgs, torch.zeros(512)), batch_size=128, shuffle=True)
    trainer = ns_gan.NSGANTrainer(model, loader, loader, loader)
    torch.manual_seed(1000 + rank)                          # ... but different noise per rank
    trainer.train(num_epochs=2, G_lr=2e-4, D_lr=2e-4, D_steps=1)
    assert len(trainer.Dlosses) == 8 and all(np.isfinite(trainer.Dlosses)) and all(np.isfinite(trainer.Glosses))
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    mine = flat.cpu() if same_gpu else flat.clone()
    allp = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allp, mine)
    assert all(torch.equal(allp[0], t) for t in allp), "Trainer replicas diverged"
    assert trainer.gradient_exchange == ("nccl" if os.environ.get("GM_DP") == "nccl" else "peer"), trainer.gradient_exchange
    return float(trainer.Dlosses[-1])


dl = trainer_check()
if rank == 0:
    print("TRAINER_DP_OK last_D_loss=%.4f" % dl)
dist.destroy_process_group()
