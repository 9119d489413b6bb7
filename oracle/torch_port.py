"""CPU baseline port — TEST / MEASUREMENT INFRASTRUCTURE, not product code.

The reference (/root/reference, pure Python on PyTorch) cannot travel to the GPU box,
so bench.py's `cpu_baseline` leg and `--impl reference` arm time THIS port of the
reference's NSGAN train step on the box's host cores.  It performs the same work the
reference performs per step, through the same library calls: nn.Linear modules,
autograd backward through the un-detached generator output in the D step, two
torch.optim.Adam optimizers, two .item() reads (src/ns_gan.py:122-156,172-216), and
optionally the per-step `next(iter(DataLoader(shuffle=True)))` batch fetch
(src/ns_gan.py:222-226).  tests/test_torch_port.py pins it against the golden
fixtures of the unmodified reference (bit-for-bit the same float32 losses).

Only bench.py and tests/ import this module.
"""
import time

import torch
import torch.nn as nn


def build_nets(x=784, h=400, z=20):
    """Generator z->h->x (ReLU, sigmoid), Discriminator x->h->1 (ReLU, sigmoid):
    src/ns_gan.py:35-60."""
    G = nn.Sequential(nn.Linear(z, h), nn.ReLU(), nn.Linear(h, x), nn.Sigmoid())
    D = nn.Sequential(nn.Linear(x, h), nn.ReLU(), nn.Linear(h, 1), nn.Sigmoid())
    return G, D


def set_weights(G, D, W):
    """W: {name: (weight, bias)} with the reference's parameter names."""
    with torch.no_grad():
        for mod, name in ((G[0], "G.linear"), (G[2], "G.generate"), (D[0], "D.linear"), (D[2], "D.discriminate")):
            mod.weight.copy_(torch.as_tensor(W[name][0]))
            mod.bias.copy_(torch.as_tensor(W[name][1]))


class NSGANStep:
    """One reference train step (D_steps=1): src/ns_gan.py:126-156."""

    def __init__(self, G, D, z_dim=20, G_lr=2e-4, D_lr=2e-4):
        self.G, self.D, self.z_dim = G, D, z_dim
        self.optG = torch.optim.Adam(G.parameters(), lr=G_lr)   # src/ns_gan.py:107-110
        self.optD = torch.optim.Adam(D.parameters(), lr=D_lr)

    def __call__(self, images, noise_fn=None):
        noise_fn = noise_fn or (lambda b, z: torch.randn(b, z))   # src/ns_gan.py:218-220
        B = images.shape[0]
        # --- D update (src/ns_gan.py:129-142)
        self.optD.zero_grad()
        fake = self.G(noise_fn(B, self.z_dim))                    # not detached, like :184
        d_loss = torch.sum(-torch.mean(torch.log(self.D(images) + 1e-8) + torch.log(1 - self.D(fake) + 1e-8)))
        d_loss.backward()
        self.optD.step()
        d_val = d_loss.item()
        # --- G update (src/ns_gan.py:148-156)
        self.optG.zero_grad()
        fake = self.G(noise_fn(B, self.z_dim))
        g_loss = -torch.mean(torch.log(self.D(fake) + 1e-8))
        g_val = g_loss.item()
        g_loss.backward()
        self.optG.step()
        return d_val, g_val


def time_cpu_steps(batch, steps, warmup, threads=None, pool=None, with_loader=False, seed=0):
    """Time `steps` train steps at `batch` on the host CPU.  Returns (images/s,
    seconds, threads).  pool: optional float32 [N,784] image pool (else synthetic
    Bernoulli(0.1307)); with_loader: fetch each batch the reference's way (fresh
    shuffling DataLoader iterator per step)."""
    if threads:
        torch.set_num_threads(threads)
    threads = torch.get_num_threads()
    torch.manual_seed(seed)
    G, D = build_nets()
    step = NSGANStep(G, D)
    if pool is None:
        g = torch.Generator().manual_seed(3435)
        pool = (torch.rand(max(batch, 4096), 784, generator=g) < 0.1307).float()
    N = pool.shape[0]
    loader = None
    if with_loader:
        ds = torch.utils.data.TensorDataset(pool.view(N, 1, 28, 28), torch.zeros(N, dtype=torch.long))
        loader = torch.utils.data.DataLoader(ds, batch_size=batch, shuffle=True)

    def fetch(i):
        if loader is not None:
            imgs, _ = next(iter(loader))                          # src/ns_gan.py:224
            return imgs.view(imgs.shape[0], -1)
        o = (i * batch) % max(N - batch + 1, 1)
        return pool[o:o + batch]
    for i in range(warmup):
        step(fetch(i))
    t0 = time.perf_counter()
    for i in range(steps):
        step(fetch(i))
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt, threads
