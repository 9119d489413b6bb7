"""Where does the WGAN-GP step (pre-activation path) differ from the bf16-point oracle?  Compares the device's stored
pre-activations, U mask and logits of the x_hat rows with the model (golden wgp case)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden")); sys.path.insert(0, os.path.join(ROOT, "generative-models_b200"))
import torch
import gm_b200
from inputs import load_case, images_from_bits, unpack_draws, gm_init_weights, GAN_SHAPES, params_dict, B
from oracle import ref_math as R

fx = load_case("gan_wgp")
x = images_from_bits(fx)
draws = unpack_draws(fx, "step1_")
z1, eps = draws[0], draws[1]
W = gm_init_weights(GAN_SHAPES, 1234)
P = params_dict(W, np.float64)
eng = gm_b200.GanEngine(784, 400, 20, max_batch=B, variant="wgp", d_out_act="relu")
eng.load(0, [W["G.linear"][0], W["G.linear"][1], W["G.generate"][0], W["G.generate"][1]])
eng.load(1, [W["D.linear"][0], W["D.linear"][1], W["D.discriminate"][0], W["D.discriminate"][1]])
Ld = eng.d_grad(torch.from_numpy(x).cuda(), noise=torch.from_numpy(z1).cuda(), aux=torch.from_numpy(eps.reshape(-1).copy()).cuda()).item()
q = R.bf16_points
e64 = eps.astype(np.float64)
gf = R.g_forward(P, z1.astype(np.float64), q=q)
fr = R.d_forward(P, x.astype(np.float64), "relu", q=q)
fg = R.d_forward(P, gf["out"], "relu", q=q)
pre = eng.debug_read("Aall", 0, 2 * B, 400).cpu().numpy().astype(np.float64)
fake = eng.debug_read("Xall", B, B, 784).cpu().numpy().astype(np.float64)
print("fake elements differing from the model:", int((fake != gf["out"]).sum()), "of", fake.size)
print("pre_r differing:", int((pre[:B] != q("a", fr["a1"])).sum()), " pre_f differing:", int((pre[B:] != q("a", fg["a1"])).sum()),
      " max abs diff", float(np.abs(pre[:B] - q("a", fr["a1"])).max()), float(np.abs(pre[B:] - q("a", fg["a1"])).max()))
a_model = e64 * q("a", fr["a1"]) + (1 - e64) * q("a", fg["a1"])
a_devpts = e64 * pre[:B] + (1 - e64) * pre[B:]
U = eng.debug_read("DHall", 2 * B, B, 400).cpu().numpy()
w2 = P["D.discriminate.weight"].reshape(-1)
m_dev = U != 0
live = m_dev.any(axis=1)          # rows whose coef is not 0 (U was scaled in place by coef after the mask was written)
print("rows with coef != 0:", int(live.sum()), "of", B)
print("mask (live rows): device vs model", int((m_dev != (a_model > 0))[live].sum()), " device vs its own stored points",
      int((m_dev != (a_devpts > 0))[live].sum()), " (w2 == 0 units:", int((w2 == 0).sum()), ")")
bad = np.argwhere((m_dev != (a_model > 0)) & live[:, None])
for r, c in bad[:10]:
    print("  row", r, "unit", c, "a_model", a_model[r, c], "a_devpts", a_devpts[r, c])
Lo, go, info = R.gan_d_step(P, "wgp", x.astype(np.float64), z1.astype(np.float64), e64, None, q=q, pre_points=(pre[:B], pre[B:]))
gD = [v.cpu().numpy() for v in eng.views(1, eng.grads[1])]
names = ["D.linear.weight", "D.linear.bias", "D.discriminate.weight", "D.discriminate.bias"]
for n, g in zip(names, gD):
    ref = go[n].reshape(g.shape)
    print(n, float(np.linalg.norm(g - ref) / max(np.linalg.norm(ref), 1e-30)))
print("loss", Ld, Lo)
