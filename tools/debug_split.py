"""debug: where does the split-mode G step lose precision (nreg=2 engines)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "generative-models_b200"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import gm_b200
from inputs import *
from oracle import ref_math as R

def nrel(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))

W = gm_init_weights(GAN_SHAPES, 1234)
P = params_dict(W, np.float64)
fx = load_case("gan_ns")
x = images_from_bits(fx)
dr = unpack_draws(fx, "step1_")
z1, z2 = dr[0], dr[-1]
for variant in ("ns", "dra", "wgp"):
    for order in ("g_only", "d_then_g", "d_then_g_twice"):
        eng = gm_b200.GanEngine(784, 400, 20, max_batch=64, variant=variant, d_out_act="relu" if variant == "wgp" else "sigmoid", precision="split")
        eng.load(0, [W["G.linear"][0], W["G.linear"][1], W["G.generate"][0], W["G.generate"][1]])
        eng.load(1, [W["D.linear"][0], W["D.linear"][1], W["D.discriminate"][0], W["D.discriminate"][1]])
        zt = torch.from_numpy(z2).cuda()
        gen0 = eng.generate(zt).cpu().numpy()
        ref = R.g_forward(P, z2.astype(np.float64))["out"]
        if order != "g_only":
            aux = None
            if variant == "wgp":
                aux = torch.rand(64, device="cuda")
            if variant == "dra":
                aux = torch.rand(64 + 64 * 784, device="cuda")
            eng.d_grad(torch.from_numpy(x).cuda(), noise=torch.from_numpy(z1).cuda(), aux=aux)
        gen1 = eng.generate(zt).cpu().numpy()
        eng.g_grad(64, noise=zt)
        if order == "d_then_g_twice":
            eng.g_grad(64, noise=zt)
        _, gg, _ = R.gan_g_step(P, variant, z2.astype(np.float64))
        names = ["G.linear.weight", "G.linear.bias", "G.generate.weight", "G.generate.bias"]
        errs = {n.split(".", 1)[1]: "%.1e" % nrel(g.cpu().numpy(), gg[n]) for n, g in zip(names, eng.views(0, eng.grads[0]))}
        print(variant, order, "generate before/after d_grad: %.1e %.1e" % (nrel(gen0, ref), nrel(gen1, ref)), errs, flush=True)
