""" (DRAGAN) Deep regret analytic GAN — drop-in for src/dra_gan.py.
NS loss + 10 E[(||grad D(xhat)||_2 - 1)^2] around perturbed REAL data,
xhat = delta x + (1-delta)(x + C std(x) u) (src/dra_gan.py:195-220); unlike WGAN-GP the
critic keeps its sigmoid, so the penalty also back-propagates through D(xhat) itself.
"""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import numpy as np  # noqa: F401

from utils import *  # noqa: F401,F403  (to_var, to_cuda, get_data — src/utils.py)
from gm_b200.gan_api import builtin_step, Generator, Discriminator, GANBase, GANTrainerBase, G_NET, D_NET
from gm_b200 import AdamHP



class DRAGAN(GANBase):
    """ Container for D and G (src/dra_gan.py:63-74) """


class DRAGANTrainer(GANTrainerBase):
    variant = "dra"

    def train(self, num_epochs, G_lr=1e-4, D_lr=1e-4, D_steps=5):
        super().train(num_epochs, G_lr=G_lr, D_lr=D_lr, D_steps=D_steps)

    @builtin_step
    def train_D(self, images, LAMBDA=10, K=1, C=1):
        return super().train_D(images, gp_lambda=float(LAMBDA), gp_k=float(K), dra_c=float(C))

    def _draw_aux(self, images):
        gen = getattr(self, "_noise_gen", None)
        delta = torch.rand(images.shape[0], 1, generator=gen)                    # src/dra_gan.py:200
        u = torch.rand(images.shape[0], images.shape[1], generator=gen)           # src/dra_gan.py:205
        return to_cuda(torch.cat([delta.reshape(-1), u.reshape(-1)])).contiguous()


if __name__ == "__main__":
    train_iter, val_iter, test_iter = get_data()
    model = DRAGAN(image_size=784, hidden_dim=400, z_dim=20)
    trainer = DRAGANTrainer(model=model, train_iter=train_iter, val_iter=val_iter, test_iter=test_iter, viz=False)
    trainer.train(num_epochs=65, G_lr=1e-4, D_lr=1e-4, D_steps=1)
