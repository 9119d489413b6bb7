"""debug: where does the split-mode G step lose precision?"""
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
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))

W = gm_init_weights(GAN_SHAPES, 1234)
P = params_dict(W, np.float64)
fx = load_case("gan_ns")
dr = unpack_draws(fx, "step1_")
z2 = dr[-1]
B = 64
eng = gm_b200.GanEngine(784, 400, 20, max_batch=B, variant="ns", precision="split")
eng.load(0, [W["G.linear"][0], W["G.linear"][1], W["G.generate"][0], W["G.generate"][1]])
eng.load(1, [W["D.linear"][0], W["D.linear"][1], W["D.discriminate"][0], W["D.discriminate"][1]])
eng.g_grad(B, noise=torch.from_numpy(z2).cuda())
z = z2.astype(np.float64)
gf = R.g_forward(P, z)
fg = R.d_forward(P, gf["out"])
L, ddg = R.g_loss("ns", fg["d"])
ds = R.d_out_grad(fg, ddg, "sigmoid")
W1 = P["D.linear.weight"]; w2 = P["D.discriminate.weight"]
M = w2 * (fg["h"] > 0)
dx = ds * (M @ W1)
da2 = dx * gf["out"] * (1 - gf["out"])
dhg = (da2 @ P["G.generate.weight"]) * (gf["h"] > 0)
rd = lambda n, r0, rows, cols: eng.debug_read(n, r0, rows, cols).cpu().numpy()
print("Zb   ", nrel(rd("Zb", 0, B, 20), z))
print("Hg   ", nrel(rd("Hg", 0, B, 400), gf["h"]), "ones col", rd("Hg", 0, B, 416)[:, 400:403].mean(0))
print("fake ", nrel(rd("Xall", B, B, 784), gf["out"]))
print("M    ", nrel(rd("Aall", B, B, 400), M))
mh = eng.debug_read("Aall", B, B, 400, plane=1).cpu().numpy(); ml = eng.debug_read("Aall", B, B, 400, plane=2).cpu().numpy()
bad = np.argwhere(np.abs(mh.astype(np.float64) + ml - M) > 1e-4 * np.abs(M).max())
print("M bad entries:", len(bad), "rows", sorted(set(bad[:, 0].tolist()))[:40], "cols (first 30)", sorted(set(bad[:, 1].tolist()))[:30])
for r, c in bad[:6]:
    print("  row %d col %d: hi %.6g lo %.6g want %.6g  a1 %.4g" % (r, c, mh[r, c], ml[r, c], M[r, c], fg["a1"][r, c]))
print("M hi-only err", nrel(mh, M), " |lo| max", float(np.abs(ml).max()))
print("ds   ", nrel(eng.scores(B).cpu().numpy(), fg["d"].ravel()))
print("DA2  ", nrel(rd("DA2", 0, B, 784), da2))
print("DHg  ", nrel(rd("DHg", 0, B, 400), dhg))
d = rd("DA2", 0, B, 784)
err = np.abs(d - da2) / (np.abs(da2).max())
print("DA2 err by 208-col tile:", [float(err[:, c:c + 208].max()) for c in range(0, 784, 208)], "by 16 rows:", [float(err[r:r + 16].max()) for r in range(0, 64, 16)])
names = ["G.linear.weight", "G.linear.bias", "G.generate.weight", "G.generate.bias"]
_, gg, _ = R.gan_g_step(P, "ns", z)
print({n: "%.1e" % nrel(g.cpu().numpy(), gg[n]) for n, g in zip(names, eng.views(0, eng.grads[0]))})
g_w2 = da2.T @ gf["h"]
print("dW2g from device DA2 x exact h:", nrel(d.T @ gf["h"], g_w2), " device grad vs (device DA2^T device Hg):",
      nrel(eng.views(0, eng.grads[0])[2].cpu().numpy(), d.astype(np.float64).T @ rd("Hg", 0, B, 400).astype(np.float64)))
