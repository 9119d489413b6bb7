"""The reference-facing class surface (generative-models_b200/*.py) on the GPU: same
names / signatures / attributes as src/*.py, the reference's own training loop runs
unchanged against it, and losses match the golden fixtures of the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from inputs import GAN_SHAPES, STEPS, B, gm_init_weights, load_case, unpack_draws, images_from_bits

pytestmark = pytest.mark.gpu


def _load(model, W):
    sd = model.state_dict()
    for k, (w, b) in W.items():
        sd[k + ".weight"] = torch.from_numpy(w.copy())
        sd[k + ".bias"] = torch.from_numpy(b.copy())
    model.load_state_dict(sd)


class _Replay:
    """compute_noise replacement feeding the reference's recorded draws."""

    def __init__(self, draws):
        self.it = iter(draws)

    def __call__(self, batch_size, z_dim):
        return torch.from_numpy(next(self.it)).cuda()


@pytest.mark.parametrize("mod,mcls,tcls,case,kw", [
    ("ns_gan", "NSGAN", "NSGANTrainer", "ns", dict(G_lr=2e-4, D_lr=2e-4, D_steps=1)),
    ("mm_gan", "MMGAN", "MMGANTrainer", "mm", dict(G_lr=2e-4, D_lr=2e-4, D_steps=1, G_init=2)),
    ("w_gan", "WGAN", "WGANTrainer", "w", dict(G_lr=5e-5, D_lr=5e-5, D_steps=2, clip=0.01)),
    ("ls_gan", "LSGAN", "LSGANTrainer", "ls", dict(G_lr=1e-4, D_lr=1e-4, D_steps=1)),
    ("ra_gan", "RaNSGAN", "RaNSGANTrainer", "ra", dict(G_lr=2e-4, D_lr=2e-4, D_steps=1)),
    ("fisher_gan", "FisherGAN", "FisherGANTrainer", "fisher", dict(G_lr=1e-4, D_lr=1e-4, D_steps=1, RHO=1e-6)),
    ("f_gan", "fGAN", "fGANTrainer", "f_pearson", dict(method="pearson", G_lr=1e-4, D_lr=1e-4, D_steps=1)),
    ("w_gp_gan", "WGPGAN", "WGPGANTrainer", "wgp", dict(G_lr=1e-4, D_lr=1e-4, D_steps=1)),
    ("dra_gan", "DRAGAN", "DRAGANTrainer", "dra", dict(G_lr=1e-4, D_lr=1e-4, D_steps=1)),
])
def test_fused_train_matches_reference_losses(mod, mcls, tcls, case, kw):
    """trainer.train(num_epochs=1, ...) exactly as the reference's __main__ calls it."""
    m = __import__(mod)
    fx = load_case("gan_" + case)
    model = getattr(m, mcls)(784, 400, 20)
    _load(model, gm_init_weights(GAN_SHAPES, 1234))
    x = torch.from_numpy(images_from_bits(fx)).view(B, 1, 28, 28)
    it = [(x, torch.zeros(B, dtype=torch.long))] * (STEPS * kw.get("D_steps", 1))
    trainer = getattr(m, tcls)(model, it, it, it, viz=False)
    replay = _Replay(unpack_draws(fx))
    trainer.compute_noise = replay
    if case == "wgp":          # the reference draws eps (and DRAGAN delta, u) right after the noise
        trainer._draw_aux = lambda images: replay(0, 0).reshape(-1).contiguous()
    if case == "dra":
        trainer._draw_aux = lambda images: torch.cat([replay(0, 0).reshape(-1), replay(0, 0).reshape(-1)]).contiguous()
    trainer.train(num_epochs=1, **kw)
    assert trainer.num_epochs == 1 and len(trainer.Dlosses) == STEPS and len(trainer.Glosses) == STEPS
    scale = lambda v: max(abs(v), 0.5)
    for a, b in zip(trainer.Dlosses, fx["D_loss"]):
        assert abs(a - b) < 2e-3 * scale(b), (trainer.Dlosses, fx["D_loss"])
    for a, b in zip(trainer.Glosses, fx["G_loss"]):
        assert abs(a - b) < 2e-3 * scale(b), (trainer.Glosses, fx["G_loss"])
    # parameters are live views of the engine's masters and moved away from the init
    w0 = gm_init_weights(GAN_SHAPES, 1234)["D.linear"][0]
    assert model.D.linear.weight.is_cuda and not np.allclose(model.D.linear.weight.detach().cpu().numpy(), w0)
    sd = model.state_dict()
    assert list(sd.keys()) == ["G.linear.weight", "G.linear.bias", "G.generate.weight", "G.generate.bias",
                               "D.linear.weight", "D.linear.bias", "D.discriminate.weight", "D.discriminate.bias"]


def test_reference_loop_with_torch_optimizers():
    """Level-(ii) drop-in: the reference's OWN loop body (src/ns_gan.py:129-156) —
    zero_grad / train_D / backward / torch.optim.Adam.step / item — against our classes."""
    import ns_gan
    fx = load_case("gan_ns")
    model = ns_gan.NSGAN(784, 400, 20)
    _load(model, gm_init_weights(GAN_SHAPES, 1234))
    x = torch.from_numpy(images_from_bits(fx)).view(B, 1, 28, 28)
    it = [(x, torch.zeros(B, dtype=torch.long))] * STEPS
    trainer = ns_gan.NSGANTrainer(model, it, it, it)
    trainer.compute_noise = _Replay(unpack_draws(fx))
    images = trainer.process_batch(trainer.train_iter)
    D_loss0 = trainer.train_D(images)            # creates the engine; parameters now live on the GPU
    G_opt = torch.optim.Adam([p for p in model.G.parameters() if p.requires_grad], lr=2e-4)
    D_opt = torch.optim.Adam([p for p in model.D.parameters() if p.requires_grad], lr=2e-4)
    Dl, Gl = [], []
    for step in range(STEPS):
        images = trainer.process_batch(trainer.train_iter)
        D_opt.zero_grad()
        D_loss = D_loss0 if step == 0 else trainer.train_D(images)
        D_loss.backward()
        D_opt.step()
        Dl.append(D_loss.item())
        G_opt.zero_grad()
        G_loss = trainer.train_G(images)
        Gl.append(G_loss.item())
        G_loss.backward()
        G_opt.step()
    np.testing.assert_allclose(Dl, fx["D_loss"], rtol=2e-3)
    np.testing.assert_allclose(Gl, fx["G_loss"], rtol=2e-3)


def test_modules_forward_and_checkpoint_roundtrip(tmp_path):
    import ns_gan
    model = ns_gan.NSGAN(784, 400, 20)
    it = [(torch.zeros(100, 1, 28, 28), torch.zeros(100, dtype=torch.long))]
    trainer = ns_gan.NSGANTrainer(model, it, it, it)
    images = trainer.process_batch(it)           # batch 100: not a multiple of the 64/128 tiles
    trainer.train_D(images)
    z = torch.randn(36, 20)
    out = model.G(z).detach()
    Wg1, bg1, Wg2, bg2 = [p.detach() for p in model.G.parameters()]
    ref = torch.sigmoid(torch.relu(z.cuda() @ Wg1.t() + bg1) @ Wg2.t() + bg2)
    assert out.shape == (36, 784) and float((out - ref).norm() / ref.norm()) < 5e-3
    d = model.D(ref).detach()
    Wd1, bd1, Wd2, bd2 = [p.detach() for p in model.D.parameters()]
    dref = torch.sigmoid(torch.relu(ref @ Wd1.t() + bd1) @ Wd2.t() + bd2)
    assert d.shape == (36, 1) and float((d - dref).abs().max()) < 2e-3
    imgs = trainer.generate_images(1, save=False)
    assert imgs.shape == (36, 28, 28)
    path = str(tmp_path / "ck.pt")
    trainer.save_model(path)
    before = out.clone()
    with torch.no_grad():
        for p in model.parameters():
            p.zero_()
    trainer.load_model(path)
    assert float((model.G(z).detach() - before).abs().max()) < 1e-6


def test_device_resident_dataset_path():
    """train() over a real DataLoader(TensorDataset): images are packed to 1 bit/pixel in HBM
    once and batches are sampled on the device (the reference's per-step DataLoader fetch)."""
    import ns_gan
    from gm_b200.gan_api import DeviceDataset
    g = torch.Generator().manual_seed(0)
    imgs = (torch.rand(1000, 1, 28, 28, generator=g) < 0.13).float()
    ds = torch.utils.data.TensorDataset(imgs, torch.zeros(1000, dtype=torch.long))
    loader = torch.utils.data.DataLoader(ds, batch_size=100, shuffle=True)
    dd = DeviceDataset.from_loader(loader)
    assert dd is not None and len(dd) == 10 and dd.bits.shape == (1000, 98)
    idx = dd.sample()
    assert idx.shape == (100,) and idx.unique().numel() == 100          # distinct rows, like a shuffled batch
    # unpack on the GPU through the staging kernel and compare with the source rows
    model = ns_gan.NSGAN(784, 400, 20)
    tr = ns_gan.NSGANTrainer(model, loader, loader, loader)
    tr.train(num_epochs=1, G_lr=2e-4, D_lr=2e-4, D_steps=1)
    assert len(tr.Dlosses) == 10 and all(np.isfinite(tr.Dlosses)) and all(np.isfinite(tr.Glosses))
    eng = tr._engine
    eng.d_grad(dd.bits, fmt="bits", gather_idx=idx, batch=100, noise=torch.randn(100, 20, device="cuda"))
    sc = eng.scores(100)
    ref = model.D(imgs.view(1000, -1)[idx.long().cpu()]).detach()
    assert float((sc - ref.view(-1)).abs().max()) < 2e-3
    # a non-binary dataset falls back to process_batch
    loader2 = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(torch.rand(200, 1, 28, 28), torch.zeros(200)), batch_size=100)
    assert DeviceDataset.from_loader(loader2) is None


def _custom_ns_trainer():
    import ns_gan

    class MyNSTrainer(ns_gan.NSGANTrainer):
        """What README.md:31 tells a user to do: the reference's train_D / train_G bodies
        (src/ns_gan.py:172-216) typed in as torch code against self.model.G / self.model.D."""

        def train_D(self, images):
            DX_score = self.model.D(images)
            noise = self.compute_noise(images.shape[0], self.model.z_dim)
            G_output = self.model.G(noise)
            DG_score = self.model.D(G_output)
            return -torch.mean(torch.log(DX_score + 1e-8) + torch.log(1 - DG_score + 1e-8))

        def train_G(self, images):
            noise = self.compute_noise(images.shape[0], self.model.z_dim)
            DG_score = self.model.D(self.model.G(noise))
            return -torch.mean(torch.log(DG_score + 1e-8))

    return ns_gan, MyNSTrainer


def test_overridden_losses_train_through_the_cuda_path():
    """README.md:31 extension mechanism: a Trainer subclass overriding train_D / train_G with a
    torch loss on the scores is honoured by train() and reproduces the reference's NSGAN run."""
    ns_gan, MyNSTrainer = _custom_ns_trainer()
    fx = load_case("gan_ns")
    model = ns_gan.NSGAN(784, 400, 20)
    _load(model, gm_init_weights(GAN_SHAPES, 1234))
    x = torch.from_numpy(images_from_bits(fx)).view(B, 1, 28, 28)
    it = [(x, torch.zeros(B, dtype=torch.long))] * STEPS
    trainer = MyNSTrainer(model, it, it, it, viz=False)
    assert trainer._has_custom_step()
    trainer.compute_noise = _Replay(unpack_draws(fx))
    from gm_b200 import _lib
    _lib.launch_count(reset=True)
    trainer.train(num_epochs=1, G_lr=2e-4, D_lr=2e-4, D_steps=1)
    assert _lib.launch_count() > 20 * STEPS          # the GEMMs / Adam ran in libgm_b200.so, not in torch
    assert len(trainer.Dlosses) == STEPS and len(trainer.Glosses) == STEPS
    for a, b in zip(trainer.Dlosses, fx["D_loss"]):
        assert abs(a - b) < 2e-3 * max(abs(b), 0.5), (trainer.Dlosses, fx["D_loss"])
    for a, b in zip(trainer.Glosses, fx["G_loss"]):
        assert abs(a - b) < 2e-3 * max(abs(b), 0.5), (trainer.Glosses, fx["G_loss"])


def test_custom_path_gradients_match_oracle():
    """One D step of the custom-loss path: every D .grad tensor against the numpy oracle's
    gradients for the same weights / images / noise (exact fp32 and bf16-point oracles)."""
    from oracle import ref_math as R
    from inputs import params_dict
    ns_gan, MyNSTrainer = _custom_ns_trainer()
    fx = load_case("gan_ns")
    W = gm_init_weights(GAN_SHAPES, 1234)
    model = ns_gan.NSGAN(784, 400, 20)
    _load(model, W)
    xb = images_from_bits(fx)
    it = [(torch.from_numpy(xb).view(B, 1, 28, 28), torch.zeros(B, dtype=torch.long))]
    trainer = MyNSTrainer(model, it, it, it, viz=False)
    draws = unpack_draws(fx, "step1_")
    trainer.compute_noise = _Replay(draws)
    images = trainer.process_batch(it)
    loss = trainer.train_D(images)
    loss.backward()
    P = params_dict(W, np.float64)
    Lo, go, _ = R.gan_d_step(P, "ns", xb.astype(np.float64), draws[0].astype(np.float64))
    _, goq, _ = R.gan_d_step(P, "ns", xb.astype(np.float64), draws[0].astype(np.float64), q=R.bf16_points)
    assert abs(float(loss.detach()) - Lo) < 1e-3 * max(abs(Lo), 1e-3)
    for name, prm in model.D.named_parameters():
        g = prm.grad.detach().cpu().numpy().astype(np.float64).ravel()
        for ref, tol in ((goq["D." + name], 1e-3), (go["D." + name], 6e-2)):
            ref = np.asarray(ref, np.float64).ravel()
            assert np.linalg.norm(g - ref) <= tol * np.linalg.norm(ref), (name, tol)
    # G's parameters also received gradients through D (the reference computes and discards them)
    assert model.G.linear.weight.grad is not None and torch.isfinite(model.G.generate.weight.grad).all()
    # ... and they are the G-step gradients of L_D w.r.t. G: check against the oracle's g_backward of dL/dfake
    # three live D forwards exceed the two row regions of an NSGAN engine -> clear error, no silent corruption
    s1 = model.D(images); model.D(images); model.D(images)
    with pytest.raises(RuntimeError, match="overwritten"):
        s1.sum().backward()


def test_reference_made_checkpoints_reproduce_the_reference_outputs():
    """A checkpoint written by the reference's save_model (src/ns_gan.py:283-285; fixture made by
    tests/golden/make_golden.py) goes through the drop-in load_model and the CUDA forward kernels reproduce
    the reference's G(z), D(x) / VAE encoder and decoder outputs stored beside it."""
    import ns_gan
    import vae as V
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    ref = np.load(os.path.join(here, "ref_nsgan_h32_z8_outputs.npz"))
    x = np.unpackbits(ref["x_bits"])[: 16 * 784].reshape(16, 784).astype(np.float32)
    it = [(torch.from_numpy(x).view(16, 1, 28, 28), torch.zeros(16, dtype=torch.long))]
    model = ns_gan.NSGAN(784, 32, 8)
    tr = ns_gan.NSGANTrainer(model, it, it, it)
    tr.load_model(os.path.join(here, "ref_nsgan_h32_z8.ckpt"))
    with torch.no_grad():
        gz = model.G(torch.from_numpy(ref["z"])).cpu().numpy()
        dx = model.D(torch.from_numpy(x)).cpu().numpy()
    assert np.linalg.norm(gz - ref["G_z"]) / np.linalg.norm(ref["G_z"]) < 5e-3          # bf16 operands
    assert np.linalg.norm(dx.ravel() - ref["D_x"].ravel()) / np.linalg.norm(ref["D_x"]) < 5e-3
    vref = np.load(os.path.join(here, "ref_vae_h32_z8_outputs.npz"))
    vmodel = V.VAE(784, 32, 8)
    vtr = V.VAETrainer(vmodel, it, it, it)
    vtr.load_model(os.path.join(here, "ref_vae_h32_z8.ckpt"))
    vtr._ensure_engine(16)
    with torch.no_grad():
        mu, lv = vmodel.encoder(torch.from_numpy(x))
        dec = vmodel.decoder(torch.from_numpy(vref["z"]))
    assert np.linalg.norm(mu.cpu().numpy() - vref["mu"]) / np.linalg.norm(vref["mu"]) < 1e-2
    assert np.linalg.norm(lv.cpu().numpy() - vref["log_var"]) / np.linalg.norm(vref["log_var"]) < 1e-2
    assert np.linalg.norm(dec.cpu().numpy() - vref["decoded"]) / np.linalg.norm(vref["decoded"]) < 5e-3
    # and the round trip: save_model of the drop-in is readable by torch.load with the same keys
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        tr.save_model(os.path.join(d, "m.ckpt"))
        assert list(torch.load(os.path.join(d, "m.ckpt")).keys()) == list(ref["keys"])


@pytest.mark.parametrize("mod_name,cls,kw", [("ns_gan", "NSGAN", {}), ("w_gp_gan", "WGPGAN", dict(D_steps=2)), ("fisher_gan", "FisherGAN", {})])
def test_trainer_cuda_graph_replay_equals_eager_launches(mod_name, cls, kw):
    """Trainer.train at the reference's batch size can replay one captured CUDA graph per outer step (GANTrainerBase.cuda_graph,
    opt-in); losses and parameters are bit-identical to launching every kernel from the host."""
    import importlib
    mod = importlib.import_module(mod_name)
    g = torch.Generator().manual_seed(0)
    imgs = (torch.rand(1000, 1, 28, 28, generator=g) < 0.13).float()
    ds = torch.utils.data.TensorDataset(imgs, torch.zeros(1000, dtype=torch.long))

    def run(use_graph):
        torch.manual_seed(21)
        loader = torch.utils.data.DataLoader(ds, batch_size=100, shuffle=True)
        model = getattr(mod, cls)(784, 400, 20)
        tr = getattr(mod, cls + "Trainer")(model, loader, loader, loader)
        tr.cuda_graph = use_graph
        tr.train(num_epochs=2, **kw)
        tr.train(num_epochs=1, **kw)                     # a second call: fresh optimizers, counters carried over
        return tr, torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
    t0, p0 = run(False)
    t1, p1 = run(True)
    assert len(t1.Dlosses) == len(t0.Dlosses) == 3 * int(np.ceil(10 / kw.get("D_steps", 1)))
    assert t0.Dlosses == t1.Dlosses and t0.Glosses == t1.Glosses
    assert torch.equal(p0, p1)
    assert (t0._step, t0._dcount) == (t1._step, t1._dcount)
