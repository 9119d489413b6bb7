"""Pins oracle/torch_port.py (the CPU baseline bench.py times) against the golden
fixtures of the unmodified reference: same float32 losses for 3 steps."""
import numpy as np
import torch

from inputs import GAN_SHAPES, STEPS, gm_init_weights, load_case, unpack_draws, images_from_bits
from oracle import torch_port as TP


def test_port_reproduces_reference_losses():
    torch.set_num_threads(1)
    fx = load_case("gan_ns")
    G, D = TP.build_nets()
    TP.set_weights(G, D, gm_init_weights(GAN_SHAPES, 1234))
    step = TP.NSGANStep(G, D)
    x = torch.from_numpy(images_from_bits(fx))
    draws = iter(unpack_draws(fx))
    Dl, Gl = [], []
    for _ in range(STEPS):
        d, g = step(x, noise_fn=lambda b, z: torch.from_numpy(next(draws)))
        Dl.append(d)
        Gl.append(g)
    np.testing.assert_allclose(Dl, fx["D_loss"], rtol=1e-6)
    np.testing.assert_allclose(Gl, fx["G_loss"], rtol=1e-6)


def test_timer_runs():
    ips, dt, th = TP.time_cpu_steps(64, steps=3, warmup=1, with_loader=True)
    assert ips > 0 and dt > 0 and th >= 1
