"""Parity of the CUDA train step (through the C ABI) with (a) the golden fixtures
made by the unmodified reference and (b) the numpy oracle on the same seeded inputs.

Arithmetic: bf16 GEMM operands (8-bit mantissa), fp32 accumulation/epilogues/Adam.
Tolerances are stated per quantity below; measured errors are dumped to
gpurun_out/parity_gan.json so they can be quoted."""
import json
import os

import numpy as np
import pytest
import torch

from inputs import (GAN_SHAPES, STEPS, B, gm_init_weights, params_dict, load_case, unpack_draws,
                    images_from_bits)
from oracle import ref_math as R

pytestmark = pytest.mark.gpu

ROW_VARIANTS = ["ns", "mm", "ls", "w", "ra", "fisher", "f_total_variation", "f_forward_kl", "f_reverse_kl",
                "f_pearson", "f_hellinger", "f_jensen_shannon"]
GP_VARIANTS = ["wgp", "dra"]
KW = {"wgp": dict(G_lr=1e-4, D_lr=1e-4), "dra": dict(G_lr=1e-4, D_lr=1e-4), "ns": dict(G_lr=2e-4, D_lr=2e-4), "mm": dict(G_lr=2e-4, D_lr=2e-4, G_init=2),
      "ls": dict(G_lr=1e-4, D_lr=1e-4), "w": dict(G_lr=5e-5, D_lr=5e-5, D_steps=2, clip=0.01),
      "ra": dict(G_lr=2e-4, D_lr=2e-4), "fisher": dict(G_lr=1e-4, D_lr=1e-4, RHO=1e-6)}
for _m in ROW_VARIANTS:
    KW.setdefault(_m, dict(G_lr=1e-4, D_lr=1e-4))

# Tolerances.  north_star: outputs within 1e-3 relative of the fp32 reference.
#  - losses and D scores (what train_D / train_G return): 1e-3 relative vs the golden
#    fixtures of the unmodified reference.
#  - gradients, two checks: (i) vs the oracle evaluated with bf16 rounding at exactly
#    the points where the CUDA path stores bf16 GEMM operands (oracle q=bf16_points):
#    1e-3 norm-relative (measured <= 1.4e-4) -> the kernels compute the reference's arithmetic; (ii) vs the
#    exact fp32/fp64 oracle: the intrinsic bf16-operand error, which at batch 64 is
#    dominated by cancellation in 64-term sums (measured 2e-3 .. 4.2e-2, largest for
#    G.linear.weight) and shrinks with batch (2e-3 .. 3e-3 at batch 4096).
TOL_LOSS, TOL_SCORE, TOL_GRAD_Q, TOL_GRAD_BF16_B64 = 1e-3, 1e-3, 1e-3, 6e-2
_REPORT = {}


def _dump():
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_gan.json", "w") as f:
        json.dump(_REPORT, f, indent=1, sort_keys=True)


def _engine(variant, batch=B):
    import gm_b200
    eng = gm_b200.GanEngine(784, 400, 20, max_batch=batch, variant=variant,
                            d_out_act="relu" if variant == "wgp" else "sigmoid")
    W = gm_init_weights(GAN_SHAPES, 1234)
    eng.load(0, [W["G.linear"][0], W["G.linear"][1], W["G.generate"][0], W["G.generate"][1]])
    eng.load(1, [W["D.linear"][0], W["D.linear"][1], W["D.discriminate"][0], W["D.discriminate"][1]])
    return eng


def _aux_from(draws_iter, case):
    """(oracle aux, device aux tensor) for the GP variants, consuming the reference's draws."""
    if case == "wgp":
        eps = next(draws_iter)
        return eps.astype(np.float64), torch.from_numpy(eps.reshape(-1).copy()).cuda()
    if case == "dra":
        delta, u = next(draws_iter), next(draws_iter)
        dev = torch.from_numpy(np.concatenate([delta.reshape(-1), u.reshape(-1)])).cuda()
        return (delta.astype(np.float64), u.astype(np.float64)), dev
    return None, None


def _nrel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("case", ROW_VARIANTS + GP_VARIANTS)
def test_step1_against_golden_and_oracle(case):
    fx = load_case("gan_" + case)
    eng = _engine(case)
    x = images_from_bits(fx)
    draws = unpack_draws(fx, "step1_")
    z1, z2 = draws[0], draws[-1]
    aux_o, aux_d = _aux_from(iter(draws[1:]), case)
    P = params_dict(gm_init_weights(GAN_SHAPES, 1234), np.float64)
    st = dict(LAMBDA=0.0, RHO=1e-6) if case == "fisher" else None
    if case == "fisher":
        eng.fisher_state(0.0, 1e-6)
    Lo, go, info = R.gan_d_step(P, case, x.astype(np.float64), z1.astype(np.float64), aux_o, st)
    stq = dict(st) if st else None
    xd = torch.from_numpy(x).cuda()
    Ld = eng.d_grad(xd, noise=torch.from_numpy(z1).cuda(), aux=aux_d).item()
    pre_pts = None
    if case == "wgp":
        # WGAN-GP never forms the x_hat rows: a_hat = eps a(x) + (1-eps) a(G(z)) from the STORED (bf16) pre-activations of
        # the real / fake rows.  The bf16-point model is evaluated at those stored values: ~0.5 % of them differ by one
        # bf16 ulp from a float64 evaluation (fp32 accumulation order, measured 120 of 25 600), which moves a_hat by
        # ~1e-3 and flips about one near-zero unit of the penalty's mask per step (1e-3 on the gradient)
        pre = eng.debug_read("Aall", 0, 2 * B, 400).cpu().numpy().astype(np.float64)
        pre_pts = (pre[:B], pre[B:])
        # ... and they ARE the model's values up to that: at most one ulp away, < 2 % of the elements
        fake_q = R.g_forward(P, z1.astype(np.float64), q=R.bf16_points)["out"]
        ref_pre = np.concatenate([R.bf16_round(R.d_forward(P, rows, "relu", q=R.bf16_points)["a1"]) for rows in (x.astype(np.float64), fake_q)])
        assert np.abs(pre - ref_pre).max() <= 2.0 ** -7 * max(1.0, float(np.abs(ref_pre).max())), float(np.abs(pre - ref_pre).max())
        assert (pre != ref_pre).mean() < 0.02, float((pre != ref_pre).mean())
    _, goq, _ = R.gan_d_step(P, case, x.astype(np.float64), z1.astype(np.float64), aux_o, stq, q=R.bf16_points, pre_points=pre_pts)
    sc = eng.scores(2 * B).cpu().numpy()
    gD = [v.cpu().numpy() for v in eng.views(1, eng.grads[1])]
    rep = {}
    rep["D_loss_vs_golden"] = abs(Ld - float(fx["step1_D_loss"])) / max(abs(float(fx["step1_D_loss"])), 1e-3)
    rep["D_loss_vs_oracle"] = abs(Ld - Lo) / max(abs(Lo), 1e-3)
    rep["DX_score"] = _nrel(sc[:B], fx["step1_score_0"])
    rep["DG_score"] = _nrel(sc[B:], fx["step1_score_1"])
    names = ["D.linear.weight", "D.linear.bias", "D.discriminate.weight", "D.discriminate.bias"]
    for nme, g in zip(names, gD):
        rep["grad_" + nme] = _nrel(g, go[nme])
        rep["gradq_" + nme] = _nrel(g, goq[nme])
    Lgo, ggo, _ = R.gan_g_step(P, case, z2.astype(np.float64))
    _, ggoq, _ = R.gan_g_step(P, case, z2.astype(np.float64), q=R.bf16_points)
    Lg = eng.g_grad(B, noise=torch.from_numpy(z2).cuda()).item()
    gG = [v.cpu().numpy() for v in eng.views(0, eng.grads[0])]
    rep["G_loss_vs_golden"] = abs(Lg - float(fx["step1_G_loss"])) / max(abs(float(fx["step1_G_loss"])), 1e-3)
    for nme, g in zip(["G.linear.weight", "G.linear.bias", "G.generate.weight", "G.generate.bias"], gG):
        rep["grad_" + nme] = _nrel(g, ggo[nme])
        rep["gradq_" + nme] = _nrel(g, ggoq[nme])
    _REPORT["step1_" + case] = rep
    _dump()
    assert rep["D_loss_vs_golden"] < TOL_LOSS and rep["G_loss_vs_golden"] < TOL_LOSS, rep
    # WGAN-GP's critic ends in ReLU: its raw outputs are O(0.05), so the same absolute
    # error as the sigmoid critics' (O(0.5) outputs) reads ~10x larger relatively
    tol_score = 3e-3 if case == "wgp" else TOL_SCORE
    assert rep["DX_score"] < tol_score and rep["DG_score"] < tol_score, rep
    for k, v in rep.items():
        if k.startswith("gradq_"):
            assert v < TOL_GRAD_Q, (k, v, rep)
        elif k.startswith("grad_"):
            assert v < TOL_GRAD_BF16_B64, (k, v, rep)


@pytest.mark.parametrize("case", ROW_VARIANTS + GP_VARIANTS)
def test_three_step_trajectory_against_golden(case):
    import gm_b200
    fx = load_case("gan_" + case)
    kw = KW[case]
    eng = _engine(case)
    x = torch.from_numpy(images_from_bits(fx)).cuda()
    raw = iter(unpack_draws(fx))
    it = (torch.from_numpy(d).cuda() for d in raw)
    hpG = gm_b200.AdamHP.make(kw["G_lr"])
    hpD = gm_b200.AdamHP.make(kw["D_lr"], clamp=kw.get("clip", 0.0) or 0.0)
    if case == "fisher":
        eng.fisher_state(0.0, kw["RHO"])
    for _ in range(kw.get("G_init", 0)):
        eng.g_grad(B, noise=next(it))
        eng.apply(0, hpG)
    Dl, Gl = [], []
    for _ in range(STEPS):
        acc = []
        for _ in range(kw.get("D_steps", 1)):
            zd = next(it)
            acc.append(eng.d_grad(x, noise=zd, aux=_aux_from(raw, case)[1]).item())
            eng.apply(1, hpD)
        Dl.append(np.mean(acc))
        Gl.append(eng.g_grad(B, noise=next(it)).item())
        eng.apply(0, hpG)
    # losses that are differences of O(1) means (WGAN, f-GAN, Fisher) can be ~0: measure
    # the error against the scale of the terms, max(|loss|, 0.5)
    rep = {"D": [abs(a - b) / max(abs(b), 0.5) for a, b in zip(Dl, fx["D_loss"])],
           "G": [abs(a - b) / max(abs(b), 0.5) for a, b in zip(Gl, fx["G_loss"])]}
    _REPORT["traj_" + case] = rep
    _dump()
    # 3 optimizer steps in: 2e-3 (bf16 weight copies; WGAN's clamp puts every weight on
    # the same bf16 rounding boundary, its worst case, measured 1.6e-3)
    assert max(rep["D"]) < 2e-3 and max(rep["G"]) < 2e-3, (rep, Dl, Gl)
    # post-Adam weights after the 3 steps against the reference's final weights (full tensor for the small
    # ones, the fixture's sampled entries for the two big matrices).  Early Adam steps are sign-like, so bf16
    # noise on tiny gradient entries moves single weights by up to 2*lr per step: the bf16-point oracle itself
    # ends 5e-4 .. 2.9e-3 (norm-relative, worst: RaNS) from the reference over all variants; the bound is 1e-2.
    names = ["G.linear.weight", "G.linear.bias", "G.generate.weight", "G.generate.bias",
             "D.linear.weight", "D.linear.bias", "D.discriminate.weight", "D.discriminate.bias"]
    finals = [v.cpu().numpy() for v in eng.views(0)] + [v.cpu().numpy() for v in eng.views(1)]
    wrep = {}
    for nme, w in zip(names, finals):
        key = "final_" + nme
        got = w.reshape(-1).astype(np.float64)
        if key in fx:
            ref = fx[key].astype(np.float64).reshape(-1)
        else:
            ref, got = fx[key + "__samp"].astype(np.float64), got[fx[key + "__idx"]]
        wrep[nme] = float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30))
    _REPORT["traj_weights_" + case] = wrep
    _dump()
    assert max(wrep.values()) < 1e-2, wrep
    if case == "fisher":
        lam, _ = eng.fisher_state()
        assert abs(lam - float(fx["final_LAMBDA"][0])) <= 1e-2 * abs(float(fx["final_LAMBDA"][0])) + 1e-12


def test_property_batch_4096_gradients_match_fp32_torch():
    """Size-independent check at a larger batch: flat D/G gradients vs a plain PyTorch
    fp32 autograd evaluation of the same NSGAN losses on the GPU."""
    Bb = 4096
    eng = _engine("ns", Bb)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = (torch.rand(Bb, 784, device="cuda", generator=g) < 0.1307).float()
    z = torch.randn(Bb, 20, device="cuda", generator=g)
    Wg1, bg1, Wg2, bg2 = [t.clone().requires_grad_() for t in eng.views(0)]
    Wd1, bd1, Wd2, bd2 = [t.clone().requires_grad_() for t in eng.views(1)]

    def Gf(zz):
        return torch.sigmoid(torch.relu(zz @ Wg1.t() + bg1) @ Wg2.t() + bg2)

    def Df(xx):
        return torch.sigmoid(torch.relu(xx @ Wd1.t() + bd1) @ Wd2.t() + bd2)
    with torch.backends.cuda.sdp_kernel() if False else torch.enable_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        Ld_ref = -torch.mean(torch.log(Df(x) + 1e-8) + torch.log(1 - Df(Gf(z)) + 1e-8))
        gd = torch.autograd.grad(Ld_ref, [Wd1, bd1, Wd2, bd2])
        Lg_ref = -torch.mean(torch.log(Df(Gf(z)) + 1e-8))
        gg = torch.autograd.grad(Lg_ref, [Wg1, bg1, Wg2, bg2])
    Ld = eng.d_grad(x, noise=z).item()
    flatD = eng.grads[1].clone()
    Lg = eng.g_grad(Bb, noise=z).item()
    flatG = eng.grads[0].clone()
    rep = {"D_loss": abs(Ld - Ld_ref.item()) / abs(Ld_ref.item()), "G_loss": abs(Lg - Lg_ref.item()) / abs(Lg_ref.item()),
           "gD": _nrel(flatD.cpu().numpy(), torch.cat([t.reshape(-1) for t in gd]).cpu().numpy()),
           "gG": _nrel(flatG.cpu().numpy(), torch.cat([t.reshape(-1) for t in gg]).cpu().numpy())}
    _REPORT["b4096_ns"] = rep
    _dump()
    assert rep["D_loss"] < TOL_LOSS and rep["G_loss"] < TOL_LOSS and rep["gD"] < 5e-3 and rep["gG"] < 5e-3, rep


def test_generate_matches_reference_forward():
    eng = _engine("ns")
    z = torch.randn(36, 20, device="cuda")
    out = eng.generate(z)
    Wg1, bg1, Wg2, bg2 = eng.views(0)
    ref = torch.sigmoid(torch.relu(z @ Wg1.t() + bg1) @ Wg2.t() + bg2)
    assert _nrel(out.cpu().numpy(), ref.cpu().numpy()) < 5e-3


def test_infogan_steps_against_golden_and_oracle():
    """InfoGAN: D and G steps are NS with a 40-wide generator input; the Q / MI step
    (src/info_gan.py:269-304) trains G and Q."""
    import gm_b200
    from inputs import INFO_SHAPES
    fx = load_case("gan_info")
    W = gm_init_weights(INFO_SHAPES, 1234)
    eng = gm_b200.InfoGanEngine(784, 400, 20, 10, 10, max_batch=B)
    eng.load(0, [W["G.linear"][0], W["G.linear"][1], W["G.generate"][0], W["G.generate"][1]])
    eng.load(1, [W["D.linear"][0], W["D.linear"][1], W["D.discriminator"][0], W["D.discriminator"][1]])
    eng.load_q([W["Q.linear"][0], W["Q.linear"][1], W["Q.inference"][0], W["Q.inference"][1]])
    x = torch.from_numpy(images_from_bits(fx)).cuda()
    d1 = unpack_draws(fx, "step1_")
    P = {k.replace("D.discriminator", "D.discriminate"): v for k, v in params_dict(W, np.float64).items()}
    Ld = eng.d_grad(x, noise=torch.from_numpy(d1[0]).cuda()).item()
    assert abs(Ld - float(fx["step1_D_loss"])) < TOL_LOSS * abs(float(fx["step1_D_loss"]))
    Lg = eng.g_grad(B, noise=torch.from_numpy(d1[1]).cuda()).item()
    assert abs(Lg - float(fx["step1_G_loss"])) < TOL_LOSS * abs(float(fx["step1_G_loss"]))
    Lq = eng.q_grad(B, torch.from_numpy(d1[2]).cuda()).item()
    assert abs(Lq - float(fx["step1_MI_loss"])) < TOL_LOSS * abs(float(fx["step1_MI_loss"])), (Lq, fx["step1_MI_loss"])
    _, gq = R.info_q_step(P, d1[2].astype(np.float64), q=R.bf16_points)
    _, gq_exact = R.info_q_step(P, d1[2].astype(np.float64))
    names = ["G.linear.weight", "G.linear.bias", "G.generate.weight", "G.generate.bias"]
    rep = {}
    for nme, g in zip(names, eng.views(0, eng.grads[0])):
        rep["gradq_" + nme] = _nrel(g.cpu().numpy(), gq[nme])
        rep["grad_" + nme] = _nrel(g.cpu().numpy(), gq_exact[nme])
    for nme, g in zip(["Q.linear.weight", "Q.linear.bias", "Q.inference.weight", "Q.inference.bias"], eng.q_views(eng.q_grads)):
        rep["gradq_" + nme] = _nrel(g.cpu().numpy(), gq[nme])
        rep["grad_" + nme] = _nrel(g.cpu().numpy(), gq_exact[nme])
    _REPORT["step1_info_q"] = rep
    _dump()
    for k, v in rep.items():
        assert v < (TOL_GRAD_Q if k.startswith("gradq_") else TOL_GRAD_BF16_B64), (k, v, rep)
    # trajectory through the drop-in module
    import info_gan
    model = info_gan.InfoGAN(784, 400, 20, 10, 10)
    sd = model.state_dict()
    for k, (w, b) in W.items():
        sd[k + ".weight"], sd[k + ".bias"] = torch.from_numpy(w.copy()), torch.from_numpy(b.copy())
    model.load_state_dict(sd)
    xi = torch.from_numpy(images_from_bits(fx)).view(B, 1, 28, 28)
    it = [(xi, torch.zeros(B, dtype=torch.long))] * STEPS
    tr = info_gan.InfoGANTrainer(model, it, it, it)
    draws = iter(unpack_draws(fx))
    tr.compute_noise = lambda *a, **k: torch.from_numpy(next(draws)).cuda()
    tr.train(num_epochs=1, G_lr=2e-4, D_lr=2e-4, D_steps=1)
    np.testing.assert_allclose(tr.Dlosses, fx["D_loss"], rtol=2e-3)
    np.testing.assert_allclose(tr.Glosses, fx["G_loss"], rtol=2e-3)
    np.testing.assert_allclose(tr.MIlosses, fx["MI_loss"], rtol=2e-3)


def test_began_steps_against_golden_and_oracle():
    """BEGAN: autoencoder discriminator, L1 losses, K control (src/be_gan.py)."""
    import gm_b200
    from inputs import BEGAN_SHAPES
    fx = load_case("gan_began")
    W = gm_init_weights(BEGAN_SHAPES, 1234)
    eng = gm_b200.GanEngine(784, 400, 20, max_batch=B, variant="began")
    eng.load(0, [W["G.linear"][0], W["G.linear"][1], W["G.generate"][0], W["G.generate"][1]])
    eng.load(1, [W["D.encoder"][0], W["D.encoder"][1], W["D.decoder"][0], W["D.decoder"][1]])
    x = torch.from_numpy(images_from_bits(fx)).cuda()
    d1 = unpack_draws(fx, "step1_")
    P = params_dict(W, np.float64)
    eng.began_init(0.3, B)
    Ld = eng.d_grad(x, noise=torch.from_numpy(d1[0]).cuda()).item()
    st = eng.began_state()
    assert abs(Ld - float(fx["step1_D_loss"])) < TOL_LOSS * abs(float(fx["step1_D_loss"])), (Ld, fx["step1_D_loss"])
    assert abs(st[3] - float(fx["step1_DX_loss"])) < TOL_LOSS * st[3] and abs(st[4] - float(fx["step1_DG_loss"])) < TOL_LOSS * st[4]
    _, gq, _, _ = R.began_d_step(P, images_from_bits(fx).astype(np.float64), d1[0].astype(np.float64), 0.3, q=R.bf16_points)
    _, ge, _, _ = R.began_d_step(P, images_from_bits(fx).astype(np.float64), d1[0].astype(np.float64), 0.3)
    rep = {}
    names = ["D.encoder.weight", "D.encoder.bias", "D.decoder.weight", "D.decoder.bias"]
    for nme, g in zip(names, eng.views(1, eng.grads[1])):
        rep["gradq_" + nme] = _nrel(g.cpu().numpy(), gq[nme])
        rep["grad_" + nme] = _nrel(g.cpu().numpy(), ge[nme])
    Lg = eng.g_grad(B, noise=torch.from_numpy(d1[1]).cuda()).item()
    assert abs(Lg - float(fx["step1_G_loss"])) < TOL_LOSS * abs(float(fx["step1_G_loss"]))
    _, ggq = R.began_g_step(P, d1[1].astype(np.float64), q=R.bf16_points)
    _, gge = R.began_g_step(P, d1[1].astype(np.float64))
    for nme, g in zip(["G.linear.weight", "G.linear.bias", "G.generate.weight", "G.generate.bias"], eng.views(0, eng.grads[0])):
        rep["gradq_" + nme] = _nrel(g.cpu().numpy(), ggq[nme])
        rep["grad_" + nme] = _nrel(g.cpu().numpy(), gge[nme])
    _REPORT["step1_began"] = rep
    _dump()
    for k, v in rep.items():
        # sign(r - x) flips where |r - x| is below the bf16 resolution of r: the L1 subgradient is
        # discontinuous, so a handful of elements differ from the bf16-point oracle too
        assert v < (2e-2 if k.startswith("gradq_") else TOL_GRAD_BF16_B64), (k, v, rep)
    # trajectory (K control included) through the drop-in module
    import be_gan
    model = be_gan.BEGAN(784, 400, 20)
    sd = model.state_dict()
    for k, (w, b) in W.items():
        sd[k + ".weight"], sd[k + ".bias"] = torch.from_numpy(w.copy()), torch.from_numpy(b.copy())
    model.load_state_dict(sd)
    xi = torch.from_numpy(images_from_bits(fx)).view(B, 1, 28, 28)
    it = [(xi, torch.zeros(B, dtype=torch.long))] * STEPS
    tr = be_gan.BEGANTrainer(model, it, it, it)
    draws = iter(unpack_draws(fx))
    tr.compute_noise = lambda *a, **k: torch.from_numpy(next(draws)).cuda()
    tr.train(num_epochs=1, G_lr=1e-4, D_lr=1e-4, D_steps=1, GAMMA=0.5, LAMBDA=1e-3, K=0.0)
    np.testing.assert_allclose(tr.Dlosses, fx["D_loss"], rtol=2e-3)
    np.testing.assert_allclose(tr.Glosses, fx["G_loss"], rtol=2e-3)


@pytest.mark.parametrize("case", ["ns", "wgp"])
def test_lazy_gradients_are_bit_identical(case):
    """gm_gan_set_lazy_grads: gathering the split-K partials inside the Adam kernel gives exactly
    the parameters, moments and flat gradients of the finalize-then-Adam path; an out-of-order
    call sequence (G gradient requested while D's is pending) materialises D's first."""
    import gm_b200
    fx = load_case("gan_" + case)
    x = torch.from_numpy(images_from_bits(fx)).cuda()
    g = torch.Generator().manual_seed(7)
    zs = [torch.randn(B, 20, generator=g).cuda() for _ in range(6)]
    aux = [torch.rand(B, generator=g).cuda() for _ in range(3)] if case == "wgp" else [None] * 3
    hp = gm_b200.AdamHP.make(2e-4)

    def run(lazy, out_of_order=False):
        eng = _engine(case)
        eng.set_lazy_grads(lazy)
        grads = []
        for t in range(3):
            eng.d_grad(x, noise=zs[2 * t], aux=aux[t], step=t)
            if out_of_order and t == 1:
                eng.g_grad(B, noise=zs[2 * t + 1], step=t)      # D's partials would be clobbered: flushed first
                eng.apply(1, hp)
                eng.apply(0, hp)
            else:
                eng.apply(1, hp)
                eng.g_grad(B, noise=zs[2 * t + 1], step=t)
                eng.apply(0, hp)
            grads.append([eng.grads[n].clone() for n in (0, 1)])
        return [eng.params[n].clone() for n in (0, 1)], [eng.exp_avg_sq[n].clone() for n in (0, 1)], grads

    p0, v0, g0 = run(False)
    p1, v1, g1 = run(True)
    for n in (0, 1):
        assert torch.equal(p0[n], p1[n]) and torch.equal(v0[n], v1[n])
        for t in range(3):
            assert torch.equal(g0[t][n], g1[t][n])
    pa, _, _ = run(False, out_of_order=True)
    pb, _, _ = run(True, out_of_order=True)
    for n in (0, 1):
        assert torch.equal(pa[n], pb[n])


def test_fused_peer_allreduce_adam_two_ranks():
    """gm_gan_apply_allreduce (gradient SUM over CUDA-IPC peer mappings fused into Adam) against
    all_reduce + gm_gan_apply, two ranks.  On a one-GPU box both ranks share cuda:0."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    if torch.cuda.device_count() < 2:
        env["GM_PEER_SAME_GPU"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "peer_allreduce_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0 and "PEER_ALLREDUCE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
    assert "SPLIT_EXCHANGE_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]          # begin / finish halves == fused
    # RaNS / Fisher / DRAGAN batch statistics are exchanged on the device: 2 ranks x B rows reproduce the
    # gradients (and Fisher's lambda update) of one process with 2B rows
    assert "GLOBAL_BATCH_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
    # NSGANTrainer.train itself under a process group: replicas stay bitwise identical
    assert "TRAINER_DP_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


@pytest.mark.parametrize("batch", [1, 37, 129, 200])
def test_ragged_batches_match_the_oracle(batch):
    """Batches that are not multiples of the 128-row / 64-deep tiles (TMA zero-fill and store clipping do
    the bounds work), down to a single sample: losses, scores and the D / G gradients of one NSGAN step
    against the bf16-point oracle.  One engine (max_batch 256) serves every size."""
    eng = _engine("ns", 256)
    rng = np.random.default_rng(100 + batch)
    x = (rng.random((batch, 784)) < 0.1307).astype(np.float32)
    z1, z2 = rng.standard_normal((batch, 20)).astype(np.float32), rng.standard_normal((batch, 20)).astype(np.float32)
    P = params_dict(gm_init_weights(GAN_SHAPES, 1234), np.float64)
    Lo, goq, info = R.gan_d_step(P, "ns", x.astype(np.float64), z1.astype(np.float64), q=R.bf16_points)
    Ld = eng.d_grad(torch.from_numpy(x).cuda(), noise=torch.from_numpy(z1).cuda()).item()
    sc = eng.scores(2 * batch).cpu().numpy()
    assert abs(Ld - Lo) < TOL_LOSS * max(abs(Lo), 1e-3)
    assert _nrel(sc[:batch], info["dx"]) < TOL_SCORE and _nrel(sc[batch:], info["dg"]) < TOL_SCORE
    names = ["D.linear.weight", "D.linear.bias", "D.discriminate.weight", "D.discriminate.bias"]
    for nme, g in zip(names, [v.cpu().numpy() for v in eng.views(1, eng.grads[1])]):
        assert _nrel(g, goq[nme]) < TOL_GRAD_Q, (nme, batch)
    Lgo, ggq, _ = R.gan_g_step(P, "ns", z2.astype(np.float64), q=R.bf16_points)
    Lg = eng.g_grad(batch, noise=torch.from_numpy(z2).cuda()).item()
    assert abs(Lg - Lgo) < TOL_LOSS * max(abs(Lgo), 1e-3)
    for nme, g in zip(["G.linear.weight", "G.linear.bias", "G.generate.weight", "G.generate.bias"],
                      [v.cpu().numpy() for v in eng.views(0, eng.grads[0])]):
        assert _nrel(g, ggq[nme]) < TOL_GRAD_Q, (nme, batch)


def test_argument_errors_are_reported_not_executed():
    """The C ABI validates on the host and returns error codes (raised as GmError): empty and oversized
    batches, missing inputs, bad slots; nothing is launched for them."""
    import gm_b200
    eng = _engine("ns", 64)
    x = torch.zeros(64, 784, device="cuda")
    n0 = gm_b200.launch_count(reset=True)
    with pytest.raises(gm_b200.GmError, match="batch"):
        eng.d_grad(x[:0], batch=0, inv_global_batch=1.0)
    with pytest.raises(gm_b200.GmError, match="batch"):
        eng.d_grad(torch.zeros(65, 784, device="cuda"))
    with pytest.raises(gm_b200.GmError, match="batch"):
        eng.g_grad(1000)
    with pytest.raises(gm_b200.GmError, match="slot"):
        eng.d_forward(7, x)
    with pytest.raises(gm_b200.GmError):
        eng.generate(torch.zeros(65, 20, device="cuda"))
    assert gm_b200.launch_count() == 0
