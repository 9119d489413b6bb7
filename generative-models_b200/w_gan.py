""" (WGAN) Wasserstein GAN — drop-in for the reference's src/w_gan.py (which cannot be
imported as shipped: src/w_gan.py:40).  L(D) = E[D(G(z))] - E[D(x)], weights clamped to
[-clip, clip] after every D step (src/w_gan.py:158,241-243); D keeps its sigmoid
(src/w_gan.py:70); Adam despite the docstring (src/w_gan.py:119-122).
"""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import numpy as np  # noqa: F401

from utils import *  # noqa: F401,F403  (to_var, to_cuda, get_data — src/utils.py)
from gm_b200.gan_api import Generator, Discriminator, GANBase, GANTrainerBase, G_NET, D_NET
from gm_b200 import AdamHP


class WGAN(GANBase):
    """ Container for D and G (src/w_gan.py:74-85) """


class WGANTrainer(GANTrainerBase):
    variant = "w"

    def train(self, num_epochs, G_lr=5e-5, D_lr=5e-5, D_steps=5, clip=0.01):
        super().train(num_epochs, G_lr=G_lr, D_lr=D_lr, D_steps=D_steps, clip=clip)

    def clip_D_weights(self, clip):
        """ src/w_gan.py:241-243 (the fused train() applies the clamp inside the Adam kernel) """
        for parameter in self.model.D.parameters():
            parameter.data.clamp_(-clip, clip)
        self._needs_sync = True
        if self._engine is not None:
            self._engine.sync_shadows(D_NET)


if __name__ == "__main__":
    train_iter, val_iter, test_iter = get_data()
    model = WGAN(image_size=784, hidden_dim=400, z_dim=20)
    trainer = WGANTrainer(model=model, train_iter=train_iter, val_iter=val_iter, test_iter=test_iter, viz=False)
    trainer.train(num_epochs=25, G_lr=5e-5, D_lr=5e-5, D_steps=5, clip=0.01)
