""" (DCGAN) NSGAN with deep-convolutional G / D — the model the reference's README recommends ("for more complex
datasets ... DCGAN", README.md:68) and lists under To-Do (README.md:96); BASELINE configs[4]: 64x64x3 images.

There is no src/dc_gan.py in the reference.  This module gives the conv model the SAME class surface as src/ns_gan.py
so the reference's driver code runs on it unchanged:

    model = DCGAN(image_size=64 * 64 * 3, hidden_dim=64, z_dim=100)
    trainer = DCGANTrainer(model, train_iter, val_iter, test_iter, viz=False)
    trainer.train(num_epochs=25, G_lr=2e-4, D_lr=2e-4, D_steps=1)

Losses, loop and optimizers are NSGAN's (src/ns_gan.py:94-216); images cross the class boundary flattened to
[B, image_size] exactly as process_batch produces them (src/ns_gan.py:222-226) and D un-flattens.  All arithmetic runs
in the sm_100a kernels behind gm_b200.DcganEngine (im2col / col2im + tcgen05 GEMMs, BatchNorm, loss, Adam); the
nn.Conv2d / nn.ConvTranspose2d / nn.BatchNorm2d members only hold the parameters in torch's layouts, so state_dict()
has the usual DCGAN keys and shapes.  Under torchrun the trainer is data-parallel: per-rank batches and noise, NCCL
all-reduce (SUM) of the flat G and D gradients before each Adam step.
"""
import numpy as np
import torch
import torch.nn as nn

from utils import *  # noqa: F401,F403
from gm_b200 import AdamHP, GmError, DcganEngine
from gm_b200 import parallel as par
from gm_b200.gan_api import to_cuda, _FusedLoss


class Generator(nn.Module):
    """ z -> 4x4 -> 8x8 -> 16x16 -> 32x32 -> 64x64 (transposed convolutions, BatchNorm + ReLU, sigmoid output) """

    def __init__(self, image_size, hidden_dim, z_dim, channels=3):
        super().__init__()
        c = [8 * hidden_dim, 4 * hidden_dim, 2 * hidden_dim, hidden_dim, channels]
        self.l1 = nn.ConvTranspose2d(z_dim, c[0], 4, 1, 0, bias=False)
        self.l2 = nn.ConvTranspose2d(c[0], c[1], 4, 2, 1, bias=False)
        self.l3 = nn.ConvTranspose2d(c[1], c[2], 4, 2, 1, bias=False)
        self.l4 = nn.ConvTranspose2d(c[2], c[3], 4, 2, 1, bias=False)
        self.l5 = nn.ConvTranspose2d(c[3], c[4], 4, 2, 1, bias=False)
        self.bn1, self.bn2, self.bn3, self.bn4 = (nn.BatchNorm2d(k) for k in c[:4])
        self._owner = None

    def forward(self, x):
        tr = self._owner
        if tr is None:
            raise GmError("Generator is not attached to a CUDA engine yet: construct the DCGANTrainer first")
        return tr._engine_synced().generate(to_cuda(x).float())


class Discriminator(nn.Module):
    """ 64x64 -> 32x32 -> 16x16 -> 8x8 -> 4x4 -> 1 (convolutions, BatchNorm + LeakyReLU(0.2), sigmoid output) """

    def __init__(self, image_size, hidden_dim, output_dim=1, channels=3):
        super().__init__()
        if output_dim != 1:
            raise GmError("only output_dim=1 discriminators are built")
        c = [hidden_dim, 2 * hidden_dim, 4 * hidden_dim, 8 * hidden_dim]
        self.l1 = nn.Conv2d(channels, c[0], 4, 2, 1, bias=False)
        self.l2 = nn.Conv2d(c[0], c[1], 4, 2, 1, bias=False)
        self.l3 = nn.Conv2d(c[1], c[2], 4, 2, 1, bias=False)
        self.l4 = nn.Conv2d(c[2], c[3], 4, 2, 1, bias=False)
        self.l5 = nn.Conv2d(c[3], 1, 4, 1, 0, bias=False)
        self.bn2, self.bn3, self.bn4 = (nn.BatchNorm2d(k) for k in c[1:])
        self._owner = None

    def forward(self, x):
        tr = self._owner
        if tr is None:
            raise GmError("Discriminator is not attached to a CUDA engine yet: construct the DCGANTrainer first")
        return tr._engine_synced().discriminate(to_cuda(x).float().reshape(x.shape[0], -1))


class DCGAN(nn.Module):
    """ Super class to contain both Discriminator (D) and Generator (G) (as src/ns_gan.py:63-74) """

    def __init__(self, image_size=64 * 64 * 3, hidden_dim=64, z_dim=100, output_dim=1, channels=3):
        super().__init__()
        if image_size != 64 * 64 * channels:
            raise GmError("the conv path is built for 64x64 images (image_size = 64*64*channels)")
        self.__dict__.update(dict(image_size=image_size, hidden_dim=hidden_dim, z_dim=z_dim, output_dim=output_dim,
                                  channels=channels))
        self.G = Generator(image_size, hidden_dim, z_dim, channels)
        self.D = Discriminator(image_size, hidden_dim, output_dim, channels)
        for m in self.modules():                                # DCGAN initialisation (Radford et al. 2015)
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.normal_(m.weight, 0.0, 0.02)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.normal_(m.weight, 1.0, 0.02)
                nn.init.zeros_(m.bias)
        self.shape = 64


class DCGANTrainer:
    """ Object to hold data iterators, train a GAN variant (surface of src/ns_gan.py:77-290) """
    variant = "ns"

    def __init__(self, model, train_iter, val_iter, test_iter, viz=False):
        self.model = model
        self.name = model.__class__.__name__
        self.train_iter, self.val_iter, self.test_iter = train_iter, val_iter, test_iter
        self.Glosses, self.Dlosses = [], []
        self.viz = viz
        self.num_epochs = 0
        self._engine = None
        self._dirty = True               # module parameters newer than the engine's
        self._step = 0
        self._seed = int(torch.initial_seed() & 0x7FFFFFFF)
        object.__setattr__(model.G, "_owner", self)
        object.__setattr__(model.D, "_owner", self)

    # ------------------------------------------------------------------ engine <-> module parameters
    def _sd(self):
        out = {}
        for tag, mod in (("G", self.model.G), ("D", self.model.D)):
            for k, v in mod.named_parameters():
                out["%s.%s" % (tag, k)] = v.detach()
        return out

    def _engine_synced(self):
        m = self.model
        if self._engine is None:
            self._engine = DcganEngine(m.hidden_dim, m.z_dim, m.channels, variant=self.variant)
            self._dirty = True
        if self._dirty:
            self._engine.load_torch_weights(self._sd())
            self._dirty = False
        return self._engine

    def _pull(self):
        """engine -> module parameters (after training; before state_dict / save_model)"""
        if self._engine is None:
            return
        tw = self._engine.torch_weights()
        with torch.no_grad():
            for tag, mod in (("G", self.model.G), ("D", self.model.D)):
                for k, v in mod.named_parameters():
                    v.copy_(tw["%s.%s" % (tag, k)].to(v.device))
                for i, bn in ((1, getattr(mod, "bn1", None)), (2, mod.bn2), (3, mod.bn3), (4, mod.bn4)):
                    run = (self._engine.run_G.get(i - 1) if tag == "G" else self._engine.run_D.get(i - 1))
                    if bn is not None and run is not None:
                        bn.running_mean.copy_(run[0].cpu())
                        bn.running_var.copy_(run[1].cpu())

    # ------------------------------------------------------------------ reference surface
    def train(self, num_epochs, G_lr=2e-4, D_lr=2e-4, D_steps=1):
        """ Trainer.train (src/ns_gan.py:94-170): same loop and logging on the fused conv step """
        import torch.distributed as dist
        eng = self._engine_synced()
        hpG, hpD = AdamHP.make(G_lr), AdamHP.make(D_lr)
        for net in (eng.G, eng.D):                                  # fresh optimizers per train() call (src/ns_gan.py:107-110)
            net.exp_avg.zero_(); net.exp_avg_sq.zero_(); net.step = 0
        world, rank = par.world_size(), par.rank_of()
        if world > 1:                                               # replicas start from rank 0's parameters
            for net in (eng.G, eng.D):
                dist.broadcast(net.params, src=0)
                net.refresh()
        seed = par.rank_seed(self._seed, rank)
        epoch_steps = int(np.ceil(len(self.train_iter) / D_steps))
        for epoch in range(1, num_epochs + 1):
            self.model.train()
            ring = torch.zeros(D_steps + 1, epoch_steps, device="cuda")
            for i in range(epoch_steps):
                for k in range(D_steps):
                    images = self.process_batch(self.train_iter)
                    n = images.shape[0]
                    inv = par.inv_global_batch(n, world)
                    ring[k, i] = eng.d_grad(eng.stage_images(images), n, inv_global_batch=inv, seed=seed, step=self._step * D_steps + k)
                    par.sum_gradients(eng.D.grads)                  # NCCL SUM of the flat D gradient (no-op on one GPU)
                    eng.apply(1, hpD)
                ring[D_steps, i] = eng.g_grad(n, inv_global_batch=inv, seed=seed, step=self._step)
                par.sum_gradients(eng.G.grads)
                eng.apply(0, hpG)
                self._step += 1
            G_losses, D_losses = ring[D_steps].tolist(), ring[:D_steps].mean(dim=0).tolist()
            self.Glosses.extend(G_losses)
            self.Dlosses.extend(D_losses)
            print("Epoch[%d/%d], G Loss: %.4f, D Loss: %.4f" % (epoch, num_epochs, np.mean(G_losses), np.mean(D_losses)))
            self.num_epochs += 1
        self._pull()

    def _loss(self, net, loss_val):
        eng = self._engine
        mod = self.model.G if net == 0 else self.model.D
        enet = eng.G if net == 0 else eng.D
        grads = []
        for k, p in mod.named_parameters():                         # engine layout -> torch layout, per tensor
            g = enet.view(k, enet.grads).detach()
            if k.startswith("l"):
                if net == 0:
                    g = g.view(4, 4, p.shape[1], p.shape[0]).permute(3, 2, 0, 1)
                else:
                    g = g[: p.shape[0]].view(p.shape[0], 4, 4, -1).permute(0, 3, 1, 2)
            grads.append(g.reshape(-1).to(p.device))
        params = [p for _, p in mod.named_parameters()]
        flat = torch.cat(grads)
        self._dirty = True                                            # the caller's optimizer will change the module parameters
        return _FusedLoss.apply(flat.detach().requires_grad_(True), loss_val.detach().to(flat.device), flat, params)

    def train_D(self, images):
        """ Run 1 step of training for discriminator (src/ns_gan.py:172-194): returns D_loss; .backward() delivers the gradients """
        images = to_cuda(images)
        eng = self._engine_synced()
        n = images.shape[0]
        noise = self.compute_noise(n, self.model.z_dim)
        loss = eng.d_grad(eng.stage_images(images.reshape(n, -1).float()), n, noise=noise.float().contiguous())
        return self._loss(1, loss.clone())

    def train_G(self, images):
        """ Run 1 step of training for generator (src/ns_gan.py:196-216) """
        eng = self._engine_synced()
        n = images.shape[0]
        noise = self.compute_noise(n, self.model.z_dim)
        loss = eng.g_grad(n, noise=noise.float().contiguous())
        return self._loss(0, loss.clone())

    def compute_noise(self, batch_size, z_dim):
        """ Compute random noise for the generator to learn to make images from (src/ns_gan.py:218-220) """
        return to_cuda(torch.randn(batch_size, z_dim))

    def process_batch(self, iterator):
        """ Generate a process batch to be input into the discriminator D (src/ns_gan.py:222-226) """
        images, _ = next(iter(iterator))
        return to_cuda(images.view(images.shape[0], -1)).float().contiguous()

    def generate_images(self, epoch, num_outputs=36, save=True):
        """ Sample a grid from G (src/ns_gan.py:228-262 without the plotting) """
        self.model.eval()
        noise = self.compute_noise(num_outputs, self.model.z_dim)
        images = self.model.G(noise)
        return images.view(num_outputs, self.model.channels, 64, 64)

    def viz_loss(self):
        print("viz_loss: matplotlib is not installed")

    def save_model(self, savepath):
        """ Save model state dictionary (src/ns_gan.py:283-285) """
        if not self._dirty:
            self._pull()
        torch.save(self.model.state_dict(), savepath)

    def load_model(self, loadpath):
        """ Load state dictionary into model (src/ns_gan.py:287-290) """
        self.model.load_state_dict(torch.load(loadpath))
        self._dirty = True


if __name__ == "__main__":
    imgs = (torch.rand(8192, 3, 64, 64) < 0.3).float()
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(imgs, torch.zeros(8192)), batch_size=256, shuffle=True)
    model = DCGAN(image_size=64 * 64 * 3, hidden_dim=64, z_dim=100)
    trainer = DCGANTrainer(model=model, train_iter=loader, val_iter=loader, test_iter=loader, viz=False)
    trainer.train(num_epochs=1, G_lr=2e-4, D_lr=2e-4, D_steps=1)
