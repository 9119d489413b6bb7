""" (MM GAN) Minimax GAN — drop-in for the reference's src/mm_gan.py.
L(G) = E[log(1 - D(G(z)))] (src/mm_gan.py:235); G is pre-trained for G_init steps
(src/mm_gan.py:121-136).
"""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import numpy as np  # noqa: F401

from utils import *  # noqa: F401,F403  (to_var, to_cuda, get_data — src/utils.py)
from gm_b200.gan_api import Generator, Discriminator, GANBase, GANTrainerBase, G_NET, D_NET
from gm_b200 import AdamHP


class MMGAN(GANBase):
    """ Container for D and G (src/mm_gan.py:66-77) """


class MMGANTrainer(GANTrainerBase):
    variant = "mm"

    def train(self, num_epochs, G_lr=2e-4, D_lr=2e-4, D_steps=1, G_init=5):
        super().train(num_epochs, G_lr=G_lr, D_lr=D_lr, D_steps=D_steps, G_init=G_init)

    def _pre_train(self, num_epochs, hpG, hpD, D_steps, extra):
        super()._pre_train(num_epochs, hpG, hpD, D_steps, extra)
        G_init = extra.get("G_init", 0)
        for _ in range(G_init):                      # src/mm_gan.py:121-136
            images = self.process_batch(self.train_iter)
            self._fused_G(images.shape[0], hpG)
        if G_init > 0:
            print("G pre-trained for %d training steps." % G_init)


if __name__ == "__main__":
    train_iter, val_iter, test_iter = get_data()
    model = MMGAN(image_size=784, hidden_dim=400, z_dim=20)
    trainer = MMGANTrainer(model=model, train_iter=train_iter, val_iter=val_iter, test_iter=test_iter, viz=False)
    trainer.train(num_epochs=25, G_lr=2e-4, D_lr=2e-4, D_steps=1, G_init=5)
