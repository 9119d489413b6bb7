"""Deterministic, torch-independent inputs shared by make_golden.py (which feeds
them to the reference) and the tests (which feed them to the oracle and to the
CUDA path).  numpy PCG64 streams are stable across platforms and versions."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
B, X, H, Z = 64, 784, 400, 20
STEPS = 3

GAN_SHAPES = [("G.linear", (H, Z)), ("G.generate", (X, H)),
              ("D.linear", (H, X)), ("D.discriminate", (1, H))]
INFO_SHAPES = [("G.linear", (H, Z + 20)), ("G.generate", (X, H)),
               ("D.linear", (H, X)), ("D.discriminator", (1, H)),
               ("Q.linear", (H, X)), ("Q.inference", (20, H))]
BEGAN_SHAPES = [("G.linear", (H, Z)), ("G.generate", (X, H)),
                ("D.encoder", (H, X)), ("D.decoder", (X, H))]
VAE_SHAPES = [("encoder.linear", (H, X)), ("encoder.mu", (Z, H)), ("encoder.log_var", (Z, H)),
              ("decoder.linear", (H, Z)), ("decoder.recon", (X, H))]


def gm_init_weights(shapes, seed):
    """U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias (nn.Linear's default
    distribution), drawn from numpy PCG64. Returns {name: (W[out,in], b[out])}."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, (o, i) in shapes:
        bound = 1.0 / np.sqrt(i)
        W = rng.uniform(-bound, bound, size=(o, i)).astype(np.float32)
        b = rng.uniform(-bound, bound, size=(o,)).astype(np.float32)
        out[name] = (W, b)
    return out


def gm_images(batch, seed=3435, p=0.1307, x=X):
    """i.i.d. Bernoulli(p) {0,1} float32 images [batch, x]."""
    rng = np.random.default_rng(seed)
    return (rng.random((batch, x)) < p).astype(np.float32)


def params_dict(weights, dtype=np.float32):
    P = {}
    for name, (W, b) in weights.items():
        P[name + ".weight"] = W.astype(dtype)
        P[name + ".bias"] = b.astype(dtype)
    return P


def load_case(name):
    return dict(np.load(os.path.join(HERE, name + ".npz")))


def unpack_draws(fx, prefix=""):
    kinds, shapes, flat = fx[prefix + "draw_kind"], fx[prefix + "draw_shape"], fx[prefix + "draws"]
    out, off = [], 0
    for k, shp in zip(kinds, shapes):
        shp = tuple(int(s) for s in shp if s > 0)
        n = int(np.prod(shp))
        out.append(flat[off:off + n].reshape(shp))
        off += n
    assert off == flat.size
    return out


def images_from_bits(fx, batch=B):
    return np.unpackbits(fx["images_bits"])[: batch * X].reshape(batch, X).astype(np.float32)


def check_summary(fx, prefix, arr, rtol, atol=0.0):
    """Compare `arr` with a fixture entry written by make_golden.summarise():
    full tensor if small, else sampled entries; plus the L2-norm relative error.
    Returns the norm-relative error of what was compared."""
    a = np.asarray(arr, dtype=np.float64).reshape(-1)
    if prefix in fx:
        ref = fx[prefix].astype(np.float64)
        got = a
    else:
        idx = fx[prefix + "__idx"]
        ref = fx[prefix + "__samp"].astype(np.float64)
        got = a[idx]
    denom = max(np.linalg.norm(ref), 1e-30)
    err = np.linalg.norm(got - ref) / denom
    l2 = float(fx[prefix + "__l2"])
    l2err = abs(np.linalg.norm(a) - l2) / max(l2, 1e-30)
    assert err <= rtol or np.linalg.norm(got - ref) <= atol, (prefix, err, rtol)
    assert l2err <= rtol or abs(np.linalg.norm(a) - l2) <= atol, (prefix, "l2", l2err, rtol)
    return err
