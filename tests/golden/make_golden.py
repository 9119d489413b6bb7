#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference (read-only at
/root/reference/src) on CPU with seeded synthetic inputs.

This script only runs in the build container (the reference does not travel to
the GPU box); its outputs, tests/golden/*.npz, are committed.  What is stored per
case (all float32 unless noted):

  images_bits   packed {0,1} batch  [B,784] -> np.packbits
  draws         every torch.randn / torch.rand tensor the reference drew, in
                call order (noise, epsilon, ...), concatenated + an index table
  D_loss,G_loss per-step losses from the reference's own Trainer.train()
  step1_*       step-1 detail from a manual replay of the loop body that calls the
                reference's own train_D/train_G + torch.optim.Adam: DX/DG scores,
                every gradient (full for small tensors; L2 norm, sum and 4096
                sampled entries for the two big matrices)
  final_*       weights after all steps (same summarisation)

Initial weights are NOT stored: they come from `gm_init_weights(seed)` below
(numpy PCG64, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) like nn.Linear's default), which
the tests re-create bit-exactly.

Usage:  python tests/golden/make_golden.py            # all cases
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
NSAMP = 4096


# --------------------------------------------------------------------------
# shims so the reference imports here: matplotlib / IPython are
# not installed; w_gan.py has a broken import (src/w_gan.py:40)
# --------------------------------------------------------------------------
def install_shims():
    for name in ["matplotlib", "matplotlib.pyplot", "IPython", "IPython.display"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.setdefault("display", lambda *a, **k: None)
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import utils as ref_utils  # the reference's utils
    pkg = types.ModuleType("src")
    pkg.utils = ref_utils
    sys.modules["src"] = pkg
    sys.modules["src.utils"] = ref_utils
    return ref_utils


def import_ref(modname):
    ref_utils = install_shims()
    mod = __import__(modname)
    if modname == "w_gan":      # bare names used after `from src import utils`
        mod.to_cuda = ref_utils.to_cuda
        mod.to_var = ref_utils.to_var
        mod.get_data = ref_utils.get_data
    return mod


sys.path.insert(0, HERE)
from inputs import (B, X, H, Z, STEPS, GAN_SHAPES, VAE_SHAPES,  # noqa: E402
                    gm_init_weights, gm_images)


def load_weights(module, weights):
    """module: nn.Module with sub-Linear layers addressed by dotted names."""
    sd = module.state_dict()
    for name, (W, b) in weights.items():
        sd[name + ".weight"] = torch.from_numpy(W.copy())
        sd[name + ".bias"] = torch.from_numpy(b.copy())
    module.load_state_dict(sd)


class DrawRecorder:
    """Record every torch.randn / torch.rand result (call order)."""

    def __init__(self):
        self.draws = []
        self._randn, self._rand = torch.randn, torch.rand

    def __enter__(self):
        def randn(*a, **k):
            t = self._randn(*a, **k)
            self.draws.append(("randn", t.detach().clone().numpy()))
            return t

        def rand(*a, **k):
            t = self._rand(*a, **k)
            self.draws.append(("rand", t.detach().clone().numpy()))
            return t
        torch.randn, torch.rand = randn, rand
        return self

    def __exit__(self, *exc):
        torch.randn, torch.rand = self._randn, self._rand


def summarise(prefix, arr, out, rng_seed=7):
    a = np.asarray(arr, dtype=np.float32).reshape(-1)
    if a.size <= 8192:
        out[prefix] = a
    else:
        idx = np.random.default_rng(rng_seed).choice(a.size, NSAMP, replace=False)
        idx.sort()
        out[prefix + "__idx"] = idx.astype(np.int64)
        out[prefix + "__samp"] = a[idx]
    out[prefix + "__l2"] = np.float64(np.sqrt(np.sum(a.astype(np.float64) ** 2)))
    out[prefix + "__sum"] = np.float64(np.sum(a.astype(np.float64)))


def pack_draws(draws, out):
    kinds, shapes, flat = [], [], []
    for kind, arr in draws:
        kinds.append(0 if kind == "randn" else 1)
        shapes.append(list(arr.shape) + [0] * (2 - arr.ndim))
        flat.append(arr.astype(np.float32).reshape(-1))
    out["draw_kind"] = np.asarray(kinds, dtype=np.int64)
    out["draw_shape"] = np.asarray(shapes, dtype=np.int64).reshape(-1, 2)
    out["draws"] = np.concatenate(flat) if flat else np.zeros(0, np.float32)


# --------------------------------------------------------------------------
# GAN cases
# --------------------------------------------------------------------------
GAN_CASES = {
    # name: (module, model cls, trainer cls, train kwargs, extra)
    "ns":      ("ns_gan", "NSGAN", "NSGANTrainer", dict(G_lr=2e-4, D_lr=2e-4, D_steps=1)),
    "mm":      ("mm_gan", "MMGAN", "MMGANTrainer", dict(G_lr=2e-4, D_lr=2e-4, D_steps=1, G_init=2)),
    "ls":      ("ls_gan", "LSGAN", "LSGANTrainer", dict(G_lr=1e-4, D_lr=1e-4, D_steps=1)),
    "w":       ("w_gan", "WGAN", "WGANTrainer", dict(G_lr=5e-5, D_lr=5e-5, D_steps=2, clip=0.01)),
    "wgp":     ("w_gp_gan", "WGPGAN", "WGPGANTrainer", dict(G_lr=1e-4, D_lr=1e-4, D_steps=1)),
    "dra":     ("dra_gan", "DRAGAN", "DRAGANTrainer", dict(G_lr=1e-4, D_lr=1e-4, D_steps=1)),
    "ra":      ("ra_gan", "RaNSGAN", "RaNSGANTrainer", dict(G_lr=2e-4, D_lr=2e-4, D_steps=1)),
    "fisher":  ("fisher_gan", "FisherGAN", "FisherGANTrainer", dict(G_lr=1e-4, D_lr=1e-4, D_steps=1, RHO=1e-6)),
}
GAN_CASES["info"] = ("info_gan", "InfoGAN", "InfoGANTrainer", dict(G_lr=2e-4, D_lr=2e-4, D_steps=1))
GAN_CASES["began"] = ("be_gan", "BEGAN", "BEGANTrainer", dict(G_lr=1e-4, D_lr=1e-4, D_steps=1, GAMMA=0.5, LAMBDA=1e-3, K=0.0))
for _m in ["total_variation", "forward_kl", "reverse_kl", "pearson", "hellinger", "jensen_shannon"]:
    GAN_CASES["f_" + _m] = ("f_gan", "fGAN", "fGANTrainer",
                            dict(method=_m, G_lr=1e-4, D_lr=1e-4, D_steps=1))


def make_gan_case(case):
    modname, mcls, tcls, kw = GAN_CASES[case]
    mod = import_ref(modname)
    torch.set_num_threads(1)
    from inputs import INFO_SHAPES, BEGAN_SHAPES
    shapes = INFO_SHAPES if case == "info" else (BEGAN_SHAPES if case == "began" else GAN_SHAPES)
    weights = gm_init_weights(shapes, seed=1234)

    def build_model():
        if case == "info":
            return getattr(mod, mcls)(X, H, Z, 10, 10)
        return getattr(mod, mcls)(X, H, Z)

    imgs = gm_images(B)
    images4d = torch.from_numpy(imgs.copy()).view(B, 1, 28, 28)
    labels = torch.zeros(B, dtype=torch.long)
    d_steps = kw.get("D_steps", 1)
    out = {"images_bits": np.packbits(imgs.astype(np.uint8), axis=None)}

    # ---- (1) the reference's own train() ---------------------------------
    model = build_model()
    load_weights(model, weights)
    # list iterator: next(iter(list)) == first batch every time, no RNG use;
    # len(list) fixes epoch_steps = ceil(len/D_steps) (src/ns_gan.py:114)
    it = [(images4d, labels)] * (STEPS * d_steps)
    trainer = getattr(mod, tcls)(model, it, it, it, viz=False)
    torch.manual_seed(20240923)
    noises = []
    if case == "info":      # structured noise z + onehot + N(0,1): record what compute_noise returns
        orig_cn = trainer.compute_noise

        def cn(*a, **k):
            t = orig_cn(*a, **k)
            noises.append(("randn", t.detach().clone().numpy()))
            return t
        trainer.compute_noise = cn
    with DrawRecorder() as rec:
        trainer.train(num_epochs=1, **kw)
    out["D_loss"] = np.asarray(trainer.Dlosses, dtype=np.float64)
    out["G_loss"] = np.asarray(trainer.Glosses, dtype=np.float64)
    if case == "info":
        out["MI_loss"] = np.asarray(trainer.MIlosses, dtype=np.float64)
    pack_draws(noises if case == "info" else rec.draws, out)
    for name, p in model.state_dict().items():
        summarise("final_" + name, p.numpy(), out)
    if case == "fisher":
        out["final_LAMBDA"] = trainer.LAMBDA.detach().numpy().astype(np.float32)

    # ---- (2) step-1 detail: reference train_D/train_G called by hand -----
    model2 = build_model()
    load_weights(model2, weights)
    tr2 = getattr(mod, tcls)(model2, it, it, it, viz=False)
    noises2 = []
    if case == "info":
        orig_cn2 = tr2.compute_noise

        def cn2(*a, **k):
            t = orig_cn2(*a, **k)
            noises2.append(("randn", t.detach().clone().numpy()))
            return t
        tr2.compute_noise = cn2
    if case.startswith("f_"):
        tr2.loss_fnc = mod.Divergence(kw["method"])
    if case == "fisher":
        from utils import to_var
        tr2.LAMBDA = to_var(torch.zeros(1))
        tr2.RHO = to_var(torch.tensor(kw["RHO"]))
    scores = []
    hook = model2.D.register_forward_hook(lambda m, i, o: scores.append(o.detach().clone().numpy()))
    torch.manual_seed(20240923)
    images = tr2.process_batch(it)
    if case == "mm":
        # G_init pre-steps consume RNG first (src/mm_gan.py:121-136); replaying the
        # detail for them is not needed: detail below is for a fresh (D,G) step on
        # the *initial* weights with its own seed, stored under step1_draws.
        pass
    with DrawRecorder() as rec2:
        D_loss = tr2.train_D(images, 0.3) if case == "began" else tr2.train_D(images)
        if case == "began":
            out["step1_DX_loss"], out["step1_DG_loss"] = np.float64(D_loss[1].item()), np.float64(D_loss[2].item())
        if isinstance(D_loss, tuple):
            D_loss = D_loss[0]
        for p in model2.parameters():
            p.grad = None
        D_loss.sum().backward()
        n_scores_d = len(scores)
        out["step1_D_loss"] = np.float64(D_loss.sum().item())
        for name, p in model2.D.named_parameters():
            summarise("step1_Dgrad_" + name, p.grad.numpy(), out)
        for p in model2.parameters():
            p.grad = None
        G_loss = tr2.train_G(images)
        G_loss.backward()
        out["step1_G_loss"] = np.float64(G_loss.item())
        for name, p in model2.G.named_parameters():
            summarise("step1_Ggrad_" + name, p.grad.numpy(), out)
        if case == "info":
            for p in model2.parameters():
                p.grad = None
            MI = tr2.train_Q(images)
            MI.backward()
            out["step1_MI_loss"] = np.float64(MI.item())
            for name, p in model2.G.named_parameters():
                summarise("step1_MI_Ggrad_" + name, p.grad.numpy(), out)
            for name, p in model2.Q.named_parameters():
                summarise("step1_MI_Qgrad_" + name, p.grad.numpy(), out)
    hook.remove()
    for k, s in enumerate(scores):
        out["step1_score_%d" % k] = s.reshape(-1)
    out["step1_n_scores_D"] = np.int64(n_scores_d)
    d2 = {}
    pack_draws(noises2 if case == "info" else rec2.draws, d2)
    for k, v in d2.items():
        out["step1_" + k] = v
    np.savez_compressed(os.path.join(HERE, "gan_%s.npz" % case), **out)
    print("%-18s D %s  G %s" % (case, np.round(out["D_loss"], 6), np.round(out["G_loss"], 6)))


# --------------------------------------------------------------------------
# VAE case
# --------------------------------------------------------------------------
def make_vae_case():
    mod = import_ref("vae")
    torch.set_num_threads(1)
    weights = gm_init_weights(VAE_SHAPES, seed=4321)
    imgs = gm_images(B)
    images4d = torch.from_numpy(imgs.copy()).view(B, 1, 28, 28)
    labels = torch.zeros(B, dtype=torch.long)
    out = {"images_bits": np.packbits(imgs.astype(np.uint8), axis=None)}
    it = [(images4d, labels)] * STEPS
    model = mod.VAE(X, H, Z)
    load_weights(model, weights)
    trainer = mod.VAETrainer(model, it, it[:1], it[:1], viz=False)
    torch.manual_seed(20240923)
    with DrawRecorder() as rec:
        trainer.train(num_epochs=1, lr=1e-3, weight_decay=1e-5)
    out["recon_loss"] = np.asarray(trainer.recon_loss, dtype=np.float64)
    out["kl_loss"] = np.asarray(trainer.kl_loss, dtype=np.float64)
    pack_draws(rec.draws, out)   # STEPS train draws + 1 validation draw
    for name, p in model.state_dict().items():
        summarise("final_" + name, p.numpy(), out)

    model2 = mod.VAE(X, H, Z)
    load_weights(model2, weights)
    tr2 = mod.VAETrainer(model2, it, it[:1], it[:1], viz=False)
    torch.manual_seed(20240923)
    with DrawRecorder() as rec2:
        recon, kl = tr2.compute_batch(it[0])
        (recon + kl).backward()
    out["step1_recon"] = np.float64(recon.item())
    out["step1_kl"] = np.float64(kl.item())
    for name, p in model2.named_parameters():
        summarise("step1_grad_" + name, p.grad.numpy(), out)
    d2 = {}
    pack_draws(rec2.draws, d2)
    for k, v in d2.items():
        out["step1_" + k] = v
    np.savez_compressed(os.path.join(HERE, "vae.npz"), **out)
    print("vae recon %s kl %s" % (np.round(out["recon_loss"], 4), np.round(out["kl_loss"], 5)))


# --------------------------------------------------------------------------
# reference-made checkpoints (src/ns_gan.py:283-290, src/vae.py save_model): written by the reference's own
# save_model, loaded by the drop-in load_model in tests/test_dropin_gpu.py.  Small models (hidden 32, z 8) keep
# the files at ~200 KB; the reference's outputs on fixed inputs are stored beside them.
# --------------------------------------------------------------------------
def make_checkpoints():
    rng = np.random.default_rng(99)
    x = gm_images(16, seed=5)
    ns = import_ref("ns_gan")
    torch.manual_seed(77)
    model = ns.NSGAN(784, 32, 8)
    it = [(torch.from_numpy(x).view(16, 1, 28, 28), torch.zeros(16, dtype=torch.long))]
    tr = ns.NSGANTrainer(model, it, it, it, viz=False)
    tr.train(num_epochs=2, G_lr=2e-4, D_lr=2e-4, D_steps=1)          # two real training steps before saving
    path = os.path.join(HERE, "ref_nsgan_h32_z8.ckpt")
    tr.save_model(path)
    z = rng.standard_normal((16, 8)).astype(np.float32)
    with torch.no_grad():
        gz = model.G(torch.from_numpy(z)).numpy()
        dx = model.D(torch.from_numpy(x)).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_nsgan_h32_z8_outputs.npz"), z=z, x_bits=np.packbits(x.astype(np.uint8)), G_z=gz, D_x=dx,
                        keys=np.array(list(model.state_dict().keys())))
    vae = import_ref("vae")
    torch.manual_seed(78)
    vmodel = vae.VAE(784, 32, 8)
    vtr = vae.VAETrainer(vmodel, it, it, it, viz=False)
    vtr.train(num_epochs=2, lr=1e-3, weight_decay=1e-5)
    vpath = os.path.join(HERE, "ref_vae_h32_z8.ckpt")
    vtr.save_model(vpath)
    with torch.no_grad():
        mu, lv = vmodel.encoder(torch.from_numpy(x))
        dec = vmodel.decoder(torch.from_numpy(z)).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_vae_h32_z8_outputs.npz"), z=z, x_bits=np.packbits(x.astype(np.uint8)), mu=mu.numpy(),
                        log_var=lv.numpy(), decoded=dec, keys=np.array(list(vmodel.state_dict().keys())))
    print("checkpoints:", os.path.getsize(path), os.path.getsize(vpath), "bytes")


def make_ae_case():
    """src/ae.py: 3 train steps of the reference's AutoencoderTrainer on one fixed batch (no randomness in the model)."""
    mod = import_ref("ae")
    rng = np.random.default_rng(4321)
    Hh = 32
    W = {"encoder.linear": (rng.uniform(-1 / 28, 1 / 28, (Hh, X)).astype(np.float32), rng.uniform(-1 / 28, 1 / 28, (Hh,)).astype(np.float32)),
         "decoder.linear": (rng.uniform(-Hh ** -0.5, Hh ** -0.5, (X, Hh)).astype(np.float32), rng.uniform(-Hh ** -0.5, Hh ** -0.5, (X,)).astype(np.float32))}
    x = gm_images(B, seed=3435)
    it = [(torch.from_numpy(x).view(B, 1, 28, 28), torch.zeros(B, dtype=torch.long))] * STEPS
    out = {"images_bits": np.packbits(x.astype(np.uint8))}
    for k, (w, b) in W.items():
        out["init_" + k + ".weight"], out["init_" + k + ".bias"] = w, b
    model = mod.Autoencoder(X, Hh)
    load_weights(model, W)
    tr = mod.AutoencoderTrainer(model, it, it[:1], it[:1], viz=False)
    tr.train(num_epochs=1, lr=1e-3, weight_decay=1e-5)
    out["recon_loss"] = np.asarray(tr.recon_loss, np.float64)
    for name, p in model.state_dict().items():
        summarise("final_" + name, p.numpy(), out)
    model2 = mod.Autoencoder(X, Hh)
    load_weights(model2, W)
    tr2 = mod.AutoencoderTrainer(model2, it, it[:1], it[:1], viz=False)
    loss = tr2.compute_batch(it[0])
    loss.backward()
    out["step1_loss"] = np.float64(loss.item())
    for name, p in model2.named_parameters():
        summarise("step1_grad_" + name, p.grad.numpy(), out)
    np.savez_compressed(os.path.join(HERE, "ae.npz"), **out)
    print("ae recon", np.round(out["recon_loss"], 3))


if __name__ == "__main__":
    which = sys.argv[1:] or (list(GAN_CASES) + ["vae", "checkpoints", "ae"])
    for c in which:
        if c == "vae":
            make_vae_case()
        elif c == "checkpoints":
            make_checkpoints()
        elif c == "ae":
            make_ae_case()
        else:
            make_gan_case(c)
