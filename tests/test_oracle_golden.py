"""Pins the numpy oracle (oracle/ref_math.py) against fixtures produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only.

Tolerances: the oracle runs in float64 on float32 inputs, the reference in
float32; differences are fp32 rounding of the reference (losses ~1e-6 rel,
gradients ~1e-5 norm-rel, 3-step weights ~1e-5 of the *update* size)."""
import numpy as np
import pytest

from inputs import (GAN_SHAPES, VAE_SHAPES, STEPS, gm_init_weights, params_dict, load_case,
                    unpack_draws, images_from_bits, check_summary)
from oracle import ref_math as R

GAN_KW = {
    "ns": dict(G_lr=2e-4, D_lr=2e-4), "mm": dict(G_lr=2e-4, D_lr=2e-4, G_init=2),
    "ls": dict(G_lr=1e-4, D_lr=1e-4), "w": dict(G_lr=5e-5, D_lr=5e-5, D_steps=2, clip=0.01),
    "wgp": dict(G_lr=1e-4, D_lr=1e-4), "dra": dict(G_lr=1e-4, D_lr=1e-4),
    "ra": dict(G_lr=2e-4, D_lr=2e-4), "fisher": dict(G_lr=1e-4, D_lr=1e-4, RHO=1e-6),
}
for _m in ["total_variation", "forward_kl", "reverse_kl", "pearson", "hellinger", "jensen_shannon"]:
    GAN_KW["f_" + _m] = dict(G_lr=1e-4, D_lr=1e-4)


@pytest.mark.parametrize("case", sorted(GAN_KW))
def test_gan_step1_detail(case):
    fx = load_case("gan_" + case)
    P = params_dict(gm_init_weights(GAN_SHAPES, 1234), np.float64)
    x = images_from_bits(fx).astype(np.float64)
    draws = [d.astype(np.float64) for d in unpack_draws(fx, "step1_")]
    it = iter(draws)
    z = next(it)
    aux = None
    if case == "wgp":
        aux = next(it)
    if case == "dra":
        aux = (next(it), next(it))
    st = dict(LAMBDA=0.0, RHO=1e-6) if case == "fisher" else None
    L, g, info = R.gan_d_step(P, case, x, z, aux, st)
    assert abs(L - fx["step1_D_loss"]) <= 2e-6 * max(1, abs(fx["step1_D_loss"]))
    np.testing.assert_allclose(info["dx"].ravel(), fx["step1_score_0"], rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(info["dg"].ravel(), fx["step1_score_1"], rtol=2e-5, atol=1e-7)
    for k, v in g.items():
        check_summary(fx, "step1_Dgrad_" + k[2:], v, rtol=3e-4, atol=1e-9)
    z2 = next(it)
    Lg, gg, _ = R.gan_g_step(P, case, z2)
    assert abs(Lg - fx["step1_G_loss"]) <= 2e-6 * max(1, abs(fx["step1_G_loss"]))
    for k, v in gg.items():
        check_summary(fx, "step1_Ggrad_" + k[2:], v, rtol=3e-4, atol=1e-9)


@pytest.mark.parametrize("case", sorted(GAN_KW))
def test_gan_three_step_trajectory(case):
    fx = load_case("gan_" + case)
    P0 = params_dict(gm_init_weights(GAN_SHAPES, 1234), np.float64)
    P = dict(P0)
    x = images_from_bits(fx).astype(np.float64)
    draws = [d.astype(np.float64) for d in unpack_draws(fx)]
    Dl, Gl, st = R.gan_train(P, case, x, draws, STEPS, **GAN_KW[case])
    np.testing.assert_allclose(Dl, fx["D_loss"], rtol=5e-5, atol=2e-6)
    np.testing.assert_allclose(Gl, fx["G_loss"], rtol=5e-5, atol=2e-6)
    for k in P:
        # compare the accumulated update, not the weight itself (a stricter test)
        if "final_" + k in fx:
            upd = P[k].ravel() - P0[k].ravel()
            ref_upd = fx["final_" + k].astype(np.float64) - P0[k].ravel()
        else:
            idx = fx["final_" + k + "__idx"]
            upd = P[k].ravel()[idx] - P0[k].ravel()[idx]
            ref_upd = fx["final_" + k + "__samp"].astype(np.float64) - P0[k].ravel()[idx]
        denom = max(np.linalg.norm(ref_upd), 1e-12)
        # fp32 weights quantise a ~1e-4-sized update to ~2^-24*|w| ~ 1e-9..6e-9
        assert np.linalg.norm(upd - ref_upd) / denom < 2e-3, (k, np.linalg.norm(upd - ref_upd) / denom)
    if case == "fisher":
        np.testing.assert_allclose(st["LAMBDA"], fx["final_LAMBDA"][0], rtol=1e-4, atol=1e-12)


def test_wgp_penalty_from_interpolated_preactivations_is_the_same_number():
    """The CUDA WGAN-GP path never forms x_hat (gp_hat_kernel): D's first layer is linear, so the hidden pre-activation of
    x_hat = eps x + (1-eps) G(z) is eps a(x) + (1-eps) a(G(z)).  The oracle's penalty fed with that pre-activation must be
    the penalty of the reference's formulation (src/w_gp_gan.py:197-215, pinned by the golden fixture) to rounding."""
    fx = load_case("gan_wgp")
    P = params_dict(gm_init_weights(GAN_SHAPES, 1234), np.float64)
    x = images_from_bits(fx).astype(np.float64)
    z, eps = [d.astype(np.float64) for d in unpack_draws(fx, "step1_")[:2]]
    fake = R.g_forward(P, z)["out"]
    xhat = eps * x + (1 - eps) * fake
    gp0, g0, i0 = R.gradient_penalty(P, xhat, "relu")
    a1 = eps * R.d_forward(P, x, "relu")["a1"] + (1 - eps) * R.d_forward(P, fake, "relu")["a1"]
    gp1, g1, i1 = R.gradient_penalty(P, xhat, "relu", a1=a1)
    assert abs(gp0 - gp1) <= 1e-12 * max(1.0, abs(gp0))
    np.testing.assert_allclose(i1["n"], i0["n"], rtol=1e-11, atol=1e-13)
    for k in g0:
        np.testing.assert_allclose(g1[k], g0[k], rtol=1e-9, atol=1e-13)
    # and the quantising model evaluated at given storage points reduces to the plain one when the points are its own
    q = R.bf16_points
    fr, fg = R.d_forward(P, x, "relu", q=q), R.d_forward(P, R.g_forward(P, z, q=q)["out"], "relu", q=q)
    La, ga, _ = R.gan_d_step(P, "wgp", x, z, eps, q=q)
    Lb, gb, _ = R.gan_d_step(P, "wgp", x, z, eps, q=q, pre_points=(q("a", fr["a1"]), q("a", fg["a1"])))
    assert La == Lb and all(np.array_equal(ga[k], gb[k]) for k in ga)


def test_gan_oracle_float32_matches_float64():
    fx = load_case("gan_ns")
    x = images_from_bits(fx)
    draws = unpack_draws(fx)
    out = {}
    for dt in (np.float32, np.float64):
        P = params_dict(gm_init_weights(GAN_SHAPES, 1234), dt)
        out[dt] = R.gan_train(P, "ns", x.astype(dt), [d.astype(dt) for d in draws], STEPS, 2e-4, 2e-4)
    np.testing.assert_allclose(out[np.float32][0], out[np.float64][0], rtol=2e-5)
    np.testing.assert_allclose(out[np.float32][1], out[np.float64][1], rtol=2e-5)


def test_vae_step1_and_trajectory():
    fx = load_case("vae")
    P0 = params_dict(gm_init_weights(VAE_SHAPES, 4321), np.float64)
    x = images_from_bits(fx).astype(np.float64)
    d1 = [d.astype(np.float64) for d in unpack_draws(fx, "step1_")]
    recon, kl, g, _ = R.vae_step(P0, x, d1[0])
    assert abs(recon - fx["step1_recon"]) <= 2e-6 * abs(fx["step1_recon"])
    assert abs(kl - fx["step1_kl"]) <= 2e-5 * abs(fx["step1_kl"])
    for k, v in g.items():
        check_summary(fx, "step1_grad_" + k, v, rtol=5e-5)
    P = dict(P0)
    draws = [d.astype(np.float64) for d in unpack_draws(fx)]
    Rl, Kl = R.vae_train(P, x, draws, STEPS)
    np.testing.assert_allclose(Rl, fx["recon_loss"], rtol=2e-5)
    np.testing.assert_allclose(Kl, fx["kl_loss"], rtol=2e-4)
    for k in P:
        if "final_" + k in fx:
            upd = P[k].ravel() - P0[k].ravel()
            ref_upd = fx["final_" + k].astype(np.float64) - P0[k].ravel()
        else:
            idx = fx["final_" + k + "__idx"]
            upd = P[k].ravel()[idx] - P0[k].ravel()[idx]
            ref_upd = fx["final_" + k + "__samp"].astype(np.float64) - P0[k].ravel()[idx]
        assert np.linalg.norm(upd - ref_upd) / max(np.linalg.norm(ref_upd), 1e-12) < 2e-3, k


def test_infogan_step1_and_q_step():
    from inputs import INFO_SHAPES
    fx = load_case("gan_info")
    P = params_dict(gm_init_weights(INFO_SHAPES, 1234), np.float64)
    Pn = {k.replace("D.discriminator", "D.discriminate"): v for k, v in P.items()}   # oracle uses the NS names
    x = images_from_bits(fx).astype(np.float64)
    d = [t.astype(np.float64) for t in unpack_draws(fx, "step1_")]
    L, g, _ = R.gan_d_step(Pn, "info", x, d[0])
    assert abs(L - fx["step1_D_loss"]) <= 2e-6 * abs(fx["step1_D_loss"])
    for k, v in g.items():
        check_summary(fx, "step1_Dgrad_" + k[2:].replace("discriminate", "discriminator"), v, rtol=3e-4, atol=1e-9)
    Lg, gg, _ = R.gan_g_step(Pn, "info", d[1])
    assert abs(Lg - fx["step1_G_loss"]) <= 2e-6 * abs(fx["step1_G_loss"])
    Lq, gq = R.info_q_step(Pn, d[2])
    assert abs(Lq - fx["step1_MI_loss"]) <= 2e-6 * abs(fx["step1_MI_loss"])
    for k, v in gq.items():
        pre = "step1_MI_Ggrad_" if k.startswith("G.") else "step1_MI_Qgrad_"
        check_summary(fx, pre + k[2:], v, rtol=3e-4, atol=1e-9)


def test_began_step1_and_trajectory():
    from inputs import BEGAN_SHAPES
    fx = load_case("gan_began")
    P0 = params_dict(gm_init_weights(BEGAN_SHAPES, 1234), np.float64)
    x = images_from_bits(fx).astype(np.float64)
    d = [t.astype(np.float64) for t in unpack_draws(fx, "step1_")]
    L, g, DX, DG = R.began_d_step(P0, x, d[0], 0.3)
    assert abs(L - fx["step1_D_loss"]) <= 2e-6 * abs(fx["step1_D_loss"])
    assert abs(DX - fx["step1_DX_loss"]) <= 2e-6 * DX and abs(DG - fx["step1_DG_loss"]) <= 2e-6 * DG
    for k, v in g.items():
        check_summary(fx, "step1_Dgrad_" + k[2:], v, rtol=3e-4, atol=1e-9)
    Lg, gg = R.began_g_step(P0, d[1])
    assert abs(Lg - fx["step1_G_loss"]) <= 2e-6 * abs(fx["step1_G_loss"])
    for k, v in gg.items():
        check_summary(fx, "step1_Ggrad_" + k[2:], v, rtol=3e-4, atol=1e-9)
    P = dict(P0)
    Dl, Gl, K = R.began_train(P, x, [t.astype(np.float64) for t in unpack_draws(fx)], STEPS)
    np.testing.assert_allclose(Dl, fx["D_loss"], rtol=5e-5)
    np.testing.assert_allclose(Gl, fx["G_loss"], rtol=5e-5)
