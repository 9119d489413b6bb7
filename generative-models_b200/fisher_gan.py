""" (Fisher GAN) — drop-in for src/fisher_gan.py.  D loss with the batch second-moment
constraint (src/fisher_gan.py:214-223); the multiplier LAMBDA is device state updated
LAMBDA += RHO * dL/dLAMBDA between backward and step (src/fisher_gan.py:152-159).
"""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import numpy as np  # noqa: F401

from utils import *  # noqa: F401,F403  (to_var, to_cuda, get_data — src/utils.py)
from gm_b200.gan_api import builtin_step, Generator, Discriminator, GANBase, GANTrainerBase, G_NET, D_NET
from gm_b200 import AdamHP


class FisherGAN(GANBase):
    """ Container for D and G (src/fisher_gan.py:70-81) """


class FisherGANTrainer(GANTrainerBase):
    variant = "fisher"

    def train(self, num_epochs, G_lr=1e-4, D_lr=1e-4, D_steps=1, RHO=1e-6):
        self._rho = RHO
        self._lambda0 = 0.0
        super().train(num_epochs, G_lr=G_lr, D_lr=D_lr, D_steps=D_steps, RHO=RHO)
        lam, _ = self._engine.fisher_state()
        self.LAMBDA = torch.tensor([lam])
        self.RHO = torch.tensor(RHO)

    def _pre_train(self, num_epochs, hpG, hpD, D_steps, extra):
        super()._pre_train(num_epochs, hpG, hpD, D_steps, extra)
        self._fisher_pending = (0.0, extra.get("RHO", 1e-6))    # src/fisher_gan.py:117-118
        if self._engine is not None:
            self._engine.fisher_state(*self._fisher_pending)
            self._fisher_pending = None

    def _after_engine_created(self, eng):
        pend = getattr(self, "_fisher_pending", None)
        eng.fisher_state(*(pend if pend else (0.0, getattr(self, "_rho", 1e-6))))
        self._fisher_pending = None

    @builtin_step
    def train_D(self, images):
        """ returns (D_loss, IPM_ratio) like src/fisher_gan.py:193-229; the IPM ratio is a
        logging-only quantity (with the reference's operator-precedence quirk) and is not
        reproduced: NaN is returned in its place. """
        return super().train_D(images), float("nan")


if __name__ == "__main__":
    train_iter, val_iter, test_iter = get_data()
    model = FisherGAN(image_size=784, hidden_dim=400, z_dim=20)
    trainer = FisherGANTrainer(model=model, train_iter=train_iter, val_iter=val_iter, test_iter=test_iter, viz=False)
    trainer.train(num_epochs=25, G_lr=1e-4, D_lr=1e-4, D_steps=1, RHO=1e-6)
