""" (RaNSGAN) Relativistic average non-saturating GAN — drop-in for src/ra_gan.py.
The D loss is the one the reference CODE computes (src/ra_gan.py:204-205), not its
docstring; the G loss is plain NS (src/ra_gan.py:227).
"""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import numpy as np  # noqa: F401

from utils import *  # noqa: F401,F403  (to_var, to_cuda, get_data — src/utils.py)
from gm_b200.gan_api import Generator, Discriminator, GANBase, GANTrainerBase, G_NET, D_NET
from gm_b200 import AdamHP


class RaNSGAN(GANBase):
    """ Container for D and G (src/ra_gan.py:75-86) """


class RaNSGANTrainer(GANTrainerBase):
    variant = "ra"

    def train(self, num_epochs, G_lr=2e-4, D_lr=2e-4, D_steps=1):
        super().train(num_epochs, G_lr=G_lr, D_lr=D_lr, D_steps=D_steps)


if __name__ == "__main__":
    train_iter, val_iter, test_iter = get_data()
    model = RaNSGAN(image_size=784, hidden_dim=400, z_dim=20)
    trainer = RaNSGANTrainer(model=model, train_iter=train_iter, val_iter=val_iter, test_iter=test_iter, viz=False)
    trainer.train(num_epochs=25, G_lr=2e-4, D_lr=2e-4, D_steps=1)
