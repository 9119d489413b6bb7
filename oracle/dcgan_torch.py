"""DCGAN oracle — TEST INFRASTRUCTURE, not product code.

The reference has no convolutional model (README.md:68 recommends DCGAN, README.md:96 lists it as To-Do), so parity for
BASELINE configs[4] is "doubly unpinned" (SURVEY.md 8f): this file is OUR plain-PyTorch fp32 statement of the
architecture (torch.nn Conv2d / ConvTranspose2d / BatchNorm2d on the CPU) driven by the reference's own NSGAN
formulas — train_D: -mean(log(D(x)+1e-8) + log(1-D(G(z))+1e-8)) (src/ns_gan.py:191-192), train_G:
-mean(log(D(G(z))+1e-8)) (src/ns_gan.py:214), D(images) and D(G(z)) as separate forward calls (BatchNorm statistics per
call), autograd backward, torch.optim.Adam.  tests/test_dcgan_gpu.py compares the CUDA conv path with it."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _id(t):
    return t


def bf16_points(t):
    """Rounding model of the CUDA path: every tensor it stores as a bf16 GEMM operand / activation (weights' operand
    copies, z, conv outputs before BatchNorm, activations, the generated image) is rounded to bf16 in the forward pass;
    autograd treats the rounding as identity (the CUDA path applies the bf16-weight gradient to the fp32 master too)."""
    return t + (t.to(torch.bfloat16).float() - t).detach()


def conv_transpose_k4s2(x, w, q=_id):
    """ConvTranspose2d(k=4, s=2, p=1) written the way the CUDA path computes it: one matrix product per input pixel
    (col = x Wm^T, [(kh, kw, co)] columns) followed by the fold of the tap columns (col2im).  With q = identity this is
    F.conv_transpose2d; with q = bf16_points the tap columns are rounded to bf16 before they are summed, as on the device."""
    B, Cin, H, W = x.shape
    Cout = w.shape[1]
    xr = x.permute(0, 2, 3, 1).reshape(B * H * W, Cin)
    Wm = w.permute(2, 3, 1, 0).reshape(16 * Cout, Cin)
    col = q(xr @ Wm.t())
    cols = col.view(B, H * W, 16, Cout).permute(0, 3, 2, 1).reshape(B, Cout * 16, H * W)
    return F.fold(cols, (2 * H, 2 * W), 4, padding=1, stride=2)


class Generator(nn.Module):
    q = staticmethod(_id)          # set to bf16_points to model the CUDA path's storage rounding

    def __init__(self, hd=64, z=100, ch=3):
        super().__init__()
        c = [8 * hd, 4 * hd, 2 * hd, hd, ch]
        self.l1 = nn.ConvTranspose2d(z, c[0], 4, 1, 0, bias=False)
        self.l2 = nn.ConvTranspose2d(c[0], c[1], 4, 2, 1, bias=False)
        self.l3 = nn.ConvTranspose2d(c[1], c[2], 4, 2, 1, bias=False)
        self.l4 = nn.ConvTranspose2d(c[2], c[3], 4, 2, 1, bias=False)
        self.l5 = nn.ConvTranspose2d(c[3], c[4], 4, 2, 1, bias=False)
        self.bn1, self.bn2, self.bn3, self.bn4 = (nn.BatchNorm2d(k) for k in c[:4])

    def forward(self, z):
        q = self.q
        x = q(z).view(z.shape[0], -1, 1, 1)
        x = q(torch.relu(self.bn1(q(F.conv_transpose2d(x, q(self.l1.weight), None, 1, 0)))))
        x = q(torch.relu(self.bn2(q(conv_transpose_k4s2(x, q(self.l2.weight), q)))))
        x = q(torch.relu(self.bn3(q(conv_transpose_k4s2(x, q(self.l3.weight), q)))))
        x = q(torch.relu(self.bn4(q(conv_transpose_k4s2(x, q(self.l4.weight), q)))))
        x = q(torch.sigmoid(conv_transpose_k4s2(x, q(self.l5.weight), q)))
        return x.reshape(z.shape[0], -1)                                    # flat [B, ch*64*64] like src/ns_gan.py:46


class Discriminator(nn.Module):
    q = staticmethod(_id)

    def __init__(self, hd=64, ch=3):
        super().__init__()
        c = [hd, 2 * hd, 4 * hd, 8 * hd]
        self.ch = ch
        self.l1 = nn.Conv2d(ch, c[0], 4, 2, 1, bias=False)
        self.l2 = nn.Conv2d(c[0], c[1], 4, 2, 1, bias=False)
        self.l3 = nn.Conv2d(c[1], c[2], 4, 2, 1, bias=False)
        self.l4 = nn.Conv2d(c[2], c[3], 4, 2, 1, bias=False)
        self.l5 = nn.Conv2d(c[3], 1, 4, 1, 0, bias=False)
        self.bn2, self.bn3, self.bn4 = (nn.BatchNorm2d(k) for k in c[1:])

    def logits(self, x):
        q = self.q
        x = q(x).view(x.shape[0], self.ch, 64, 64)                         # un-flatten (src/ns_gan.py:225 flattens)
        x = q(F.leaky_relu(F.conv2d(x, q(self.l1.weight), None, 2, 1), 0.2))
        x = q(F.leaky_relu(self.bn2(q(F.conv2d(x, q(self.l2.weight), None, 2, 1))), 0.2))
        x = q(F.leaky_relu(self.bn3(q(F.conv2d(x, q(self.l3.weight), None, 2, 1))), 0.2))
        x = q(F.leaky_relu(self.bn4(q(F.conv2d(x, q(self.l4.weight), None, 2, 1))), 0.2))
        return F.conv2d(x, q(self.l5.weight), None, 1, 0).view(-1, 1)

    def forward(self, x):
        return torch.sigmoid(self.logits(x))


def load_from_engine_weights(G, D, sd):
    """sd: DcganEngine.torch_weights()."""
    with torch.no_grad():
        for tag, net in (("G", G), ("D", D)):
            for name, p in net.named_parameters():
                p.copy_(sd["%s.%s" % (tag, name)].to(p.dtype))


def d_loss(G, D, images, z):
    return -torch.mean(torch.log(D(images) + 1e-8) + torch.log(1 - D(G(z)) + 1e-8))      # src/ns_gan.py:191-192


def g_loss(G, D, z):
    return -torch.mean(torch.log(D(G(z)) + 1e-8))                                          # src/ns_gan.py:214
