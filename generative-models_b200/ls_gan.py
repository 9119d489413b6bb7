""" (LS GAN) Least-squares GAN — drop-in for the reference's src/ls_gan.py.
L(D) = 1/2 E[(D(x)-b)^2] + 1/2 E[(D(G(z))-a)^2], L(G) = 1/2 E[(D(G(z))-c)^2] with the
reference defaults a=0, b=1, c=1 (src/ls_gan.py:173,197).
"""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import numpy as np  # noqa: F401

from utils import *  # noqa: F401,F403  (to_var, to_cuda, get_data — src/utils.py)
from gm_b200.gan_api import builtin_step, Generator, Discriminator, GANBase, GANTrainerBase, G_NET, D_NET
from gm_b200 import AdamHP


class LSGAN(GANBase):
    """ Container for D and G (src/ls_gan.py:64-75) """


class LSGANTrainer(GANTrainerBase):
    variant = "ls"

    def train(self, num_epochs, G_lr=1e-4, D_lr=1e-4, D_steps=1):
        super().train(num_epochs, G_lr=G_lr, D_lr=D_lr, D_steps=D_steps)

    @builtin_step
    def train_D(self, images, a=0, b=1):
        return super().train_D(images, ls_a=float(a), ls_b=float(b))

    @builtin_step
    def train_G(self, images, c=1):
        return super().train_G(images, ls_c=float(c))


if __name__ == "__main__":
    train_iter, val_iter, test_iter = get_data()
    model = LSGAN(image_size=784, hidden_dim=400, z_dim=20)
    trainer = LSGANTrainer(model=model, train_iter=train_iter, val_iter=val_iter, test_iter=test_iter, viz=False)
    trainer.train(num_epochs=25, G_lr=1e-4, D_lr=1e-4, D_steps=1)
