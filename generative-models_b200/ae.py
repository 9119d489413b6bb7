""" (AE) Standard autoencoder — drop-in for the reference's src/ae.py.

Same classes and signatures (src/ae.py:28-240): Encoder (784 -> hidden, ReLU), Decoder (hidden -> 784, sigmoid),
Autoencoder, AutoencoderTrainer with train / compute_batch / evaluate / reconstruct_images / viz_loss / save_model /
load_model and the attributes recon_loss, best_val_loss, num_epochs.  The loss follows the reference code: the sum of
squared errors (src/ae.py:158).  Forward, loss, backward and Adam run in the sm_100a kernels behind gm_b200.AeEngine.
"""
from copy import deepcopy  # noqa: F401

import numpy as np
import torch
import torch.nn as nn

from utils import *  # noqa: F401,F403
from gm_b200 import AdamHP, AeEngine, GmError
from gm_b200.gan_api import to_cuda


class Encoder(nn.Module):
    """ Feedforward network encoder (src/ae.py:28-38) """

    def __init__(self, image_size, hidden_dim):
        super().__init__()
        self.linear = nn.Linear(image_size, hidden_dim)

    def forward(self, x):
        return _trainer_of(self, "Encoder")._ensure_engine(x.shape[0]).encode(to_cuda(x).float())


class Decoder(nn.Module):
    """ Feedforward network decoder (src/ae.py:40-50) """

    def __init__(self, hidden_dim, image_size):
        super().__init__()
        self.linear = nn.Linear(hidden_dim, image_size)

    def forward(self, encoder_output):
        return _trainer_of(self, "Decoder")._ensure_engine(encoder_output.shape[0]).decode(to_cuda(encoder_output).float())


def _trainer_of(module, what):
    tr = getattr(module, "_owner", None)
    if tr is None:
        raise GmError(what + " is not attached to a CUDA engine yet: construct the AutoencoderTrainer first (there is no eager/CPU path)")
    return tr


class Autoencoder(nn.Module):
    """ Autoencoder super class to encode then decode an image (src/ae.py:53-64) """

    def __init__(self, image_size=784, hidden_dim=32):
        super().__init__()
        self.__dict__.update(dict(image_size=image_size, hidden_dim=hidden_dim))
        self.encoder = Encoder(image_size=image_size, hidden_dim=hidden_dim)
        self.decoder = Decoder(hidden_dim=hidden_dim, image_size=image_size)

    def forward(self, x):
        tr = _trainer_of(self.encoder, "Autoencoder")
        out, _ = tr._ensure_engine(x.shape[0]).forward(to_cuda(x).float())
        return out


class _AeLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, val, holder):
        ctx.holder = holder
        return val.clone()

    @staticmethod
    def backward(ctx, gout):
        for p, g in ctx.holder():
            p.grad = g * gout if p.grad is None else p.grad + g * gout
        return None, None, None


class AutoencoderTrainer:
    def __init__(self, model, train_iter, val_iter, test_iter, viz=False):
        """ Object to hold data iterators, train the model (src/ae.py:67-82) """
        self.model = model
        self.name = model.__class__.__name__
        self.train_iter, self.val_iter, self.test_iter = train_iter, val_iter, test_iter
        self.best_val_loss = 1e10
        self.debugging_image, _ = next(iter(test_iter))
        self.viz = viz
        self.recon_loss = []
        self.num_epochs = 0
        self._engine, self._max_batch = None, 0
        for mod in (model.encoder, model.decoder):
            object.__setattr__(mod, "_owner", self)

    def _ensure_engine(self, batch):
        if self._engine is not None and batch <= self._max_batch:
            self._engine.sync_all()
            return self._engine
        m, old = self.model, self._engine
        batch = max(batch, self._max_batch, 64)
        eng = AeEngine(m.image_size, m.hidden_dim, max_batch=batch)
        named = dict(m.named_parameters())
        eng.load({k: v.data for k, v in named.items()})
        views = eng.views()
        for k, p in named.items():
            p.data = views[k]                      # nn.Parameter storage == the engine's fp32 master weights
        if old is not None:
            eng.exp_avg.copy_(old.exp_avg)
            eng.exp_avg_sq.copy_(old.exp_avg_sq)
            eng.steps = old.steps
        self._engine, self._max_batch = eng, batch
        return eng

    def train(self, num_epochs, lr=1e-3, weight_decay=1e-5):
        """ Train the autoencoder (src/ae.py:84-145): a true epoch over train_iter, losses read back once per epoch """
        hp = AdamHP.make(lr, weight_decay=weight_decay)
        if self._engine is not None:
            self._engine.reset_optimizer()
        for epoch in range(1, num_epochs + 1):
            self.model.train()
            per_step = []
            for batch in self.train_iter:
                images = self._images(batch)
                eng = self._ensure_engine(images.shape[0])
                per_step.append(eng.grad(images).clone())
                eng.apply(hp)
            epoch_loss = torch.stack(per_step).tolist()
            self.recon_loss.extend(epoch_loss)
            self.model.eval()
            val_loss = self.evaluate(self.val_iter)
            if val_loss < self.best_val_loss:
                self.best_model = deepcopy(self.model.state_dict())
                self.best_val_loss = val_loss
            print("Epoch[%d/%d], Train Loss: %.4f, Val Loss: %.4f" % (epoch, num_epochs, np.mean(epoch_loss), val_loss))
            self.num_epochs += 1

    def _images(self, batch):
        images, _ = batch
        return to_cuda(images.view(images.shape[0], -1)).float().contiguous()

    def compute_batch(self, batch):
        """ Compute loss for a batch of examples (src/ae.py:147-160): .backward() delivers the gradients """
        images = self._images(batch)
        eng = self._ensure_engine(images.shape[0])
        loss = eng.grad(images).clone()
        named = dict(self.model.named_parameters())
        gviews = eng.views(eng.grads)
        return _AeLoss.apply(loss.detach().requires_grad_(True), loss, lambda: [(named[k], gviews[k]) for k in named])

    def evaluate(self, iterator):
        """ Evaluate on a given dataset (src/ae.py:162-164) """
        vals = []
        for batch in iterator:
            images = self._images(batch)
            _, loss = self._ensure_engine(images.shape[0]).forward(images, want_loss=True)
            vals.append(loss)
        return float(torch.stack(vals).mean().item())

    def reconstruct_images(self, images, epoch, save=True):
        """ Reconstruct a fixed input (src/ae.py:166-195 without the plotting) """
        batch = to_cuda(images.view(images.shape[0], -1))
        return self.model(batch).view(images.shape).squeeze()

    def viz_loss(self):
        print("viz_loss: matplotlib is not installed")

    def save_model(self, savepath):
        """ Save model state dictionary (src/ae.py:213-215) """
        torch.save(self.model.state_dict(), savepath)

    def load_model(self, loadpath):
        """ Load state dictionary into model (src/ae.py:217-220) """
        self.model.load_state_dict(torch.load(loadpath))
        if self._engine is not None:
            self._engine.sync_all()


if __name__ == "__main__":
    train_iter, val_iter, test_iter = get_data()  # noqa: F405
    model = Autoencoder(image_size=784, hidden_dim=32)
    trainer = AutoencoderTrainer(model=model, train_iter=train_iter, val_iter=val_iter, test_iter=test_iter, viz=False)
    trainer.train(num_epochs=5, lr=1e-3, weight_decay=1e-5)
