""" (f-GAN) GANs trained with an f-divergence — drop-in for src/f_gan.py.  The six
divergences of src/f_gan.py:99-142 are six compile-time-free epilogue choices of the
loss kernel; `Divergence` is kept (torch ops) for user code that calls it directly.
"""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import numpy as np  # noqa: F401

from utils import *  # noqa: F401,F403  (to_var, to_cuda, get_data — src/utils.py)
from gm_b200.gan_api import Generator, Discriminator, GANBase, GANTrainerBase, G_NET, D_NET
from gm_b200 import AdamHP


class fGAN(GANBase):
    """ Container for D and G (src/f_gan.py:71-82) """


class Divergence:
    """ f-divergence losses (src/f_gan.py:85-142), plain torch for direct use. """
    METHODS = ['total_variation', 'forward_kl', 'reverse_kl', 'pearson', 'hellinger', 'jensen_shannon']

    def __init__(self, method):
        self.method = method.lower().strip()
        assert self.method in self.METHODS, 'Invalid divergence.'

    def D_loss(self, DX_score, DG_score):
        m = self.method
        if m == 'total_variation':
            return -(torch.mean(0.5 * torch.tanh(DX_score)) - torch.mean(0.5 * torch.tanh(DG_score)))
        if m == 'forward_kl':
            return -(torch.mean(DX_score) - torch.mean(torch.exp(DG_score - 1)))
        if m == 'reverse_kl':
            return -(torch.mean(-torch.exp(DX_score)) - torch.mean(-1 - DG_score))
        if m == 'pearson':
            return -(torch.mean(DX_score) - torch.mean(0.25 * DG_score ** 2 + DG_score))
        if m == 'hellinger':
            return -(torch.mean(1 - torch.exp(DX_score)) - torch.mean((1 - torch.exp(DG_score)) / torch.exp(DG_score)))
        return -(torch.mean(2. - (1 + torch.exp(-DX_score))) - torch.mean(-(2. - torch.exp(DG_score))))

    def G_loss(self, DG_score):
        m = self.method
        if m == 'total_variation':
            return -torch.mean(0.5 * torch.tanh(DG_score))
        if m == 'forward_kl':
            return -torch.mean(torch.exp(DG_score - 1))
        if m == 'reverse_kl':
            return -torch.mean(-1 - DG_score)
        if m == 'pearson':
            return -torch.mean(0.25 * DG_score ** 2 + DG_score)
        if m == 'hellinger':
            return -torch.mean((1 - torch.exp(DG_score)) / torch.exp(DG_score))
        return -torch.mean(-(2. - torch.exp(DG_score)))


class fGANTrainer(GANTrainerBase):
    variant = "f_jensen_shannon"

    def train(self, num_epochs, method, G_lr=1e-4, D_lr=1e-4, D_steps=1):
        self._set_method(method)
        super().train(num_epochs, G_lr=G_lr, D_lr=D_lr, D_steps=D_steps)

    def _set_method(self, method):
        self.loss_fnc = Divergence(method)               # src/f_gan.py:175
        self.variant = "f_" + self.loss_fnc.method       # the engine is rebuilt if this changed


if __name__ == "__main__":
    train_iter, val_iter, test_iter = get_data()
    model = fGAN(image_size=784, hidden_dim=400, z_dim=20)
    trainer = fGANTrainer(model=model, train_iter=train_iter, val_iter=val_iter, test_iter=test_iter, viz=False)
    trainer.train(num_epochs=25, method='jensen_shannon', G_lr=1e-4, D_lr=1e-4, D_steps=1)
