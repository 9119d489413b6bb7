""" (NS GAN) Non-saturating GAN — drop-in for the reference's src/ns_gan.py.

Same classes, signatures and defaults (src/ns_gan.py:35-314); the train step runs in
the hand-written sm_100a kernels of libgm_b200.so (see gm_b200/gan_api.py).
L(D) = -E[log D(x)] - E[log(1 - D(G(z)))],  L(G) = -E[log D(G(z))].
"""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import numpy as np  # noqa: F401

from utils import *  # noqa: F401,F403  (to_var, to_cuda, get_data — src/utils.py)
from gm_b200.gan_api import Generator, Discriminator, GANBase, GANTrainerBase, G_NET, D_NET
from gm_b200 import AdamHP


class NSGAN(GANBase):
    """ Super class to contain both Discriminator (D) and Generator (G) (src/ns_gan.py:63-74) """


class NSGANTrainer(GANTrainerBase):
    """ Object to hold data iterators, train a GAN variant (src/ns_gan.py:77-290) """
    variant = "ns"

    def train(self, num_epochs, G_lr=2e-4, D_lr=2e-4, D_steps=1):
        super().train(num_epochs, G_lr=G_lr, D_lr=D_lr, D_steps=D_steps)


if __name__ == "__main__":
    train_iter, val_iter, test_iter = get_data()
    model = NSGAN(image_size=784, hidden_dim=400, z_dim=20)
    trainer = NSGANTrainer(model=model, train_iter=train_iter, val_iter=val_iter, test_iter=test_iter, viz=False)
    trainer.train(num_epochs=25, G_lr=2e-4, D_lr=2e-4, D_steps=1)
