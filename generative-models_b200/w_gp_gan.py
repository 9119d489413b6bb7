""" (WGAN-GP) Wasserstein GAN with gradient penalty — drop-in for src/w_gp_gan.py.
L(D) = E[D(G(z))] - E[D(x)] + 10 E[(||grad_xhat D(xhat)||_2 - 1)^2], xhat = eps x + (1-eps) G(z)
(src/w_gp_gan.py:197-218).  The critic ends in ReLU (src/w_gp_gan.py:61).  The penalty's
double backward is closed-form GEMMs on the same tensor-core kernel.
"""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import numpy as np  # noqa: F401

from utils import *  # noqa: F401,F403  (to_var, to_cuda, get_data — src/utils.py)
from gm_b200.gan_api import builtin_step, Generator, Discriminator, GANBase, GANTrainerBase, G_NET, D_NET
from gm_b200 import AdamHP



class WGPGAN(GANBase):
    """ Container for D and G (src/w_gp_gan.py:65-76) """


class WGPGANTrainer(GANTrainerBase):
    variant = "wgp"
    d_out_act = "relu"

    def train(self, num_epochs, G_lr=1e-4, D_lr=1e-4, D_steps=5):
        super().train(num_epochs, G_lr=G_lr, D_lr=D_lr, D_steps=D_steps)

    @builtin_step
    def train_D(self, images, LAMBDA=10):
        return super().train_D(images, gp_lambda=float(LAMBDA))

    def _draw_aux(self, images):
        return to_cuda(torch.rand(images.shape[0], 1, generator=getattr(self, "_noise_gen", None))).reshape(-1).contiguous()     # src/w_gp_gan.py:197


# the reference notebook 04 uses these names
WGANGP, WGANGPTrainer = WGPGAN, WGPGANTrainer

if __name__ == "__main__":
    train_iter, val_iter, test_iter = get_data()
    model = WGPGAN(image_size=784, hidden_dim=400, z_dim=20)
    trainer = WGPGANTrainer(model=model, train_iter=train_iter, val_iter=val_iter, test_iter=test_iter, viz=False)
    trainer.train(num_epochs=25, G_lr=1e-4, D_lr=1e-4, D_steps=1)
