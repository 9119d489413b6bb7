"""DcganEngine — NSGAN train step with DCGAN convolutional G / D (BASELINE configs[4]: 64x64x3 images, batch 8192
per GPU, data-parallel gradient all-reduce) on the sm_100a kernels of libgm_b200.so.

The reference has no convolutional model: its README recommends DCGAN (README.md:68) and lists it as To-Do
(README.md:96).  What is kept from the reference is the Trainer contract (src/ns_gan.py:94-216): train_D / train_G
with the non-saturating losses, D(images) and D(G(z)) as separate forward calls, Adam on G and on D, flattened
[B, image_size] images at the class boundary (src/ns_gan.py:222-226).  The architecture is the DCGAN of the paper the
README cites (Radford et al. 2015) with the reference's sigmoid outputs:

  G: z -> ConvT(z, 8h, 4, 1, 0) BN ReLU -> ConvT(8h, 4h, 4, 2, 1) BN ReLU -> ConvT(4h, 2h) BN ReLU -> ConvT(2h, h) BN ReLU
       -> ConvT(h, 3, 4, 2, 1) -> sigmoid                                   [B, 64, 64, 3]
  D: Conv(3, h, 4, 2, 1) LeakyReLU(0.2) -> Conv(h, 2h) BN LReLU -> Conv(2h, 4h) BN LReLU -> Conv(4h, 8h) BN LReLU
       -> Conv(8h, 1, 4, 1, 0) -> sigmoid                                   [B, 1]

Everything on the device is NHWC bf16 as row-major matrices [B*H*W, C]: a convolution is gm_im2col_k4s2 + one tcgen05
GEMM (gm_gemm_bf16), a transposed convolution one GEMM + gm_col2im_k4s2, BatchNorm / activations are gm_bn_* over the
same matrices, the loss is the MLP path's loss kernel on the conv D's logits (gm_loss_rows), Adam is gm_adam_step.
This module only sequences those C-ABI calls and owns the buffers (host language of the reference: Python).
Weights are kept in GEMM layout — conv [Cout, (kh, kw, ci)], transposed conv [(kh, kw, co), Cin] — and converted to /
from torch's Conv2d / ConvTranspose2d layouts at the state_dict boundary (torch_weights / load_torch_weights).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GmError, VARIANTS, check, lib, _ptr, _stream, gemm_bf16, adam_step

SLOPE = 0.2
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
ACT_NONE, ACT_RELU, ACT_LRELU = 0, 1, 2
C2I_NONE, C2I_SIGMOID, C2I_LRELU_GRAD, C2I_SIGMOID_GRAD = 0, 1, 2, 3


def _im2col(x, B, H, W, Cc, col):
    h = _lib.ctx()
    check(h, lib().gm_im2col_k4s2(h, _ptr(x), B, H, W, Cc, x.stride(0), _ptr(col), col.stride(0), _stream()))


def _col2im(col, B, Hi, Wi, Cc, y, mode=C2I_NONE, aux=None):
    h = _lib.ctx()
    check(h, lib().gm_col2im_k4s2(h, _ptr(col), col.stride(0), B, Hi, Wi, Cc, _ptr(y), y.stride(0), mode, _ptr(aux),
                                  aux.stride(0) if aux is not None else 0, SLOPE, _stream()))


def _bn_fwd(x, gamma, beta, act, y, stats, running):
    h = _lib.ctx()
    check(h, lib().gm_bn_forward(h, _ptr(x), x.shape[0], x.shape[1], x.stride(0), _ptr(gamma), _ptr(beta), BN_EPS, act, SLOPE,
                                 _ptr(y), y.stride(0), _ptr(stats), _ptr(running), BN_MOMENTUM, _stream()))


def _bn_bwd(dy, x, stats, gamma, beta, act, dx, dgb):
    h = _lib.ctx()
    check(h, lib().gm_bn_backward(h, _ptr(dy), _ptr(x), x.shape[0], x.shape[1], x.stride(0), _ptr(stats), _ptr(gamma), _ptr(beta),
                                  act, SLOPE, _ptr(dx), dx.stride(0), _ptr(dgb), _stream()))


class _Net:
    """Flat fp32 master parameters of one network + gradient / Adam state + bf16 GEMM operand copies."""

    def __init__(self, shapes, device):
        self.names = [n for n, _ in shapes]
        self.shapes = dict(shapes)
        self.offsets, off = {}, 0
        for n, shp in shapes:
            cnt = 1
            for s in shp:
                cnt *= s
            self.offsets[n] = (off, cnt)
            off += (cnt + 3) // 4 * 4                         # 16-byte aligned sub-tensors
        self.total = off
        kw = dict(device=device, dtype=torch.float32)
        self.params, self.grads = torch.zeros(off, **kw), torch.zeros(off, **kw)
        self.exp_avg, self.exp_avg_sq = torch.zeros(off, **kw), torch.zeros(off, **kw)
        self.step = 0
        self.bf, self.bf_t = {}, {}
        pad8 = lambda v: (v + 7) // 8 * 8                  # noqa: E731  (TMA rows are 16-byte multiples)
        for n, shp in shapes:
            if len(shp) == 2:
                self.bf[n] = torch.zeros(shp[0], pad8(shp[1]), device=device, dtype=torch.bfloat16)[:, :shp[1]]
                self.bf_t[n] = torch.zeros(shp[1], pad8(shp[0]), device=device, dtype=torch.bfloat16)[:, :shp[0]]

    def view(self, n, flat=None):
        off, cnt = self.offsets[n]
        return (self.params if flat is None else flat)[off:off + cnt].view(self.shapes[n])

    def refresh(self):
        h = _lib.ctx()
        for n in self.bf:
            w = self.view(n)
            check(h, lib().gm_cast_bf16(h, _ptr(w), w.shape[0], w.shape[1], _ptr(self.bf[n]), self.bf[n].stride(0),
                                        _ptr(self.bf_t[n]), self.bf_t[n].stride(0), _stream()))

    def adam(self, hp):
        self.step += 1
        adam_step(self.params, self.grads, self.exp_avg, self.exp_avg_sq, hp, self.step)
        self.refresh()


class DcganEngine:
    """One DCGAN (64x64xchannels images) on one GPU; see the module docstring."""

    def __init__(self, hidden_dim=64, z_dim=100, channels=3, variant="ns", device=None):
        if not torch.cuda.is_available():
            raise GmError("gm_b200 needs a CUDA (B200) device; there is no CPU fallback")
        if hidden_dim % 16 or hidden_dim <= 0:
            raise GmError("hidden_dim (the base channel width) must be a positive multiple of 16")
        if variant not in ("ns", "mm", "w", "ls") and not variant.startswith("f_"):
            raise GmError("the conv path supports the row-wise losses (ns, mm, w, ls, f_*)")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.h = _lib.ctx(self.device.index)
        self.hd, self.z, self.ch, self.variant = hidden_dim, z_dim, channels, variant
        self.zp = (z_dim + 1 + 7) // 8 * 8                       # noise rows: [z | 1 | pad], 16-byte rows
        hd = hidden_dim
        self.gc = [8 * hd, 4 * hd, 2 * hd, hd, channels]        # generator channels after each layer
        self.dc = [hd, 2 * hd, 4 * hd, 8 * hd]                  # discriminator channels after conv 1..4
        g_shapes = [("l1.weight", (16 * self.gc[0], z_dim))]
        for i in range(1, 5):
            g_shapes.append(("l%d.weight" % (i + 1), (16 * self.gc[i], self.gc[i - 1])))
        for i in range(4):
            g_shapes += [("bn%d.weight" % (i + 1), (self.gc[i],)), ("bn%d.bias" % (i + 1), (self.gc[i],))]
        d_shapes = [("l1.weight", (self.dc[0], 16 * channels))]
        for i in range(1, 4):
            d_shapes.append(("l%d.weight" % (i + 1), (self.dc[i], 16 * self.dc[i - 1])))
        d_shapes.append(("l5.weight", (16, 16 * self.dc[3])))   # 1 real output channel (row 0), padded to the MMA's N = 16
        for i in range(1, 4):
            d_shapes += [("bn%d.weight" % (i + 1), (self.dc[i],)), ("bn%d.bias" % (i + 1), (self.dc[i],))]
        self.G, self.D = _Net(g_shapes, self.device), _Net(d_shapes, self.device)
        self.run_G = {i: torch.zeros(2, self.gc[i], device=self.device) for i in range(4)}
        self.run_D = {i: torch.zeros(2, self.dc[i], device=self.device) for i in range(1, 4)}
        for r in list(self.run_G.values()) + list(self.run_D.values()):
            r[1].fill_(1.0)
        self.loss_buf = torch.zeros(2, device=self.device)
        self._bufs = {}
        self.init_weights()

    # ------------------------------------------------------------------ parameters
    def init_weights(self, seed=1234):
        """DCGAN initialisation (N(0, 0.02) conv weights, N(1, 0.02) BN scale, zero BN shift)."""
        g = torch.Generator().manual_seed(seed)
        for net in (self.G, self.D):
            for n in net.names:
                v = net.view(n)
                if n.endswith("bias"):
                    v.zero_()
                elif n.startswith("bn"):
                    v.copy_(1.0 + 0.02 * torch.randn(v.shape, generator=g))
                else:
                    v.copy_(0.02 * torch.randn(v.shape, generator=g))
        self.D.view("l5.weight")[1:].zero_()                     # padding rows of the 1-channel output layer
        self.G.refresh()
        self.D.refresh()

    def torch_weights(self):
        """{torch-style name: tensor in torch's Conv2d / ConvTranspose2d layout} (CPU fp32)."""
        out = {}
        for i in range(5):
            w = self.G.view("l%d.weight" % (i + 1)).detach().cpu()
            cout, cin = w.shape[0] // 16, w.shape[1]
            out["G.l%d.weight" % (i + 1)] = w.view(4, 4, cout, cin).permute(3, 2, 0, 1).contiguous()    # [Cin, Cout, kh, kw]
        for i in range(5):
            w = self.D.view("l%d.weight" % (i + 1)).detach().cpu()
            if i == 4:
                w = w[:1]
            cout, cin = w.shape[0], w.shape[1] // 16
            out["D.l%d.weight" % (i + 1)] = w.view(cout, 4, 4, cin).permute(0, 3, 1, 2).contiguous()    # [Cout, Cin, kh, kw]
        for net, tag in ((self.G, "G"), (self.D, "D")):
            for n in net.names:
                if n.startswith("bn"):
                    out["%s.%s" % (tag, n)] = net.view(n).detach().cpu().clone()
        return out

    def load_torch_weights(self, sd):
        for i in range(5):
            w = sd["G.l%d.weight" % (i + 1)].float()
            self.G.view("l%d.weight" % (i + 1)).copy_(w.permute(2, 3, 1, 0).reshape(-1, w.shape[0]))
            w = sd["D.l%d.weight" % (i + 1)].float()
            v = self.D.view("l%d.weight" % (i + 1))
            v[: w.shape[0]].copy_(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))
        for net, tag in ((self.G, "G"), (self.D, "D")):
            for n in net.names:
                if n.startswith("bn"):
                    net.view(n).copy_(sd["%s.%s" % (tag, n)].float())
        self.G.refresh()
        self.D.refresh()

    def _buf(self, key, rows, cols, dtype=torch.bfloat16):
        t = self._bufs.get(key)
        if t is None or t.shape[0] < rows or t.shape[1] != cols:
            t = torch.empty(rows, cols, device=self.device, dtype=dtype)
            self._bufs[key] = t
        return t[:rows]

    # ------------------------------------------------------------------ generator
    def g_forward(self, n, noise=None, seed=0, stream_id=0, tag="g"):
        """G(z) for n samples -> (images [n*4096, ch] NHWC bf16, saved activations)."""
        zb = self._buf(tag + "z", n, self.zp)
        check(self.h, lib().gm_noise_rows(self.h, _ptr(noise), _ptr(zb), n, self.z, self.zp, int(seed), int(stream_id), _stream()))
        sv = {"z": zb, "n": n}
        gc = self.gc
        c = self._buf(tag + "c0", n, 16 * gc[0])
        gemm_bf16(zb, self.G.bf["l1.weight"], c, "nt", K=self.z)                           # [n, (kh,kw,co)] = NHWC [n*16, 8h]
        x = c.view(n * 16, gc[0])
        hw = 4
        for i in range(4):
            a = self._buf(tag + "a%d" % i, x.shape[0], gc[i])
            st = self._buf(tag + "st%d" % i, 2, gc[i], torch.float32)
            _bn_fwd(x, self.G.view("bn%d.weight" % (i + 1)), self.G.view("bn%d.bias" % (i + 1)), ACT_RELU, a, st, self.run_G[i])
            sv["c%d" % i], sv["a%d" % i], sv["st%d" % i] = x, a, st
            col = self._buf(tag + "col%d" % i, a.shape[0], 16 * gc[i + 1])
            gemm_bf16(a, self.G.bf["l%d.weight" % (i + 2)], col, "nt")
            y = self._buf(tag + "c%d" % (i + 1), n * 4 * hw * hw, gc[i + 1])
            _col2im(col, n, hw, hw, gc[i + 1], y, C2I_SIGMOID if i == 3 else C2I_NONE)
            x, hw = y, 2 * hw
        sv["img"] = x
        return x, sv

    def g_backward(self, sv, dpre):
        """dpre [n*4096, ch] = dL/d(pre-sigmoid output) -> flat G gradient (self.G.grads)."""
        n, gc, G = sv["n"], self.gc, self.G
        G.grads.zero_()
        d, hw = dpre, 64
        for i in range(3, -1, -1):
            hw //= 2                                                                      # input grid of transposed conv i+2
            a, c, st = sv["a%d" % i], sv["c%d" % i], sv["st%d" % i]
            dcol = self._buf("gdcol%d" % i, a.shape[0], 16 * gc[i + 1])
            _im2col(d, n, 2 * hw, 2 * hw, gc[i + 1], dcol)                                # gradient of col2im
            gemm_bf16(dcol, a, G.view("l%d.weight" % (i + 2), G.grads), "tn")             # [16 Cout, Cin] = dcol^T a
            da = self._buf("gda%d" % i, a.shape[0], gc[i])
            gemm_bf16(dcol, G.bf_t["l%d.weight" % (i + 2)], da, "nt")                     # da = dcol Wm
            dc = self._buf("gdc%d" % i, a.shape[0], gc[i])
            dgb = self._buf("gdgb%d" % i, 2, gc[i], torch.float32)
            _bn_bwd(da, c, st, G.view("bn%d.weight" % (i + 1)), G.view("bn%d.bias" % (i + 1)), ACT_RELU, dc, dgb)
            G.view("bn%d.bias" % (i + 1), G.grads).copy_(dgb[0])
            G.view("bn%d.weight" % (i + 1), G.grads).copy_(dgb[1])
            d = dc
        gemm_bf16(d.view(n, 16 * gc[0]), sv["z"], G.view("l1.weight", G.grads), "tn", N=self.z)   # [(kh,kw,co), z]
        return G.grads

    # ------------------------------------------------------------------ discriminator
    def d_forward(self, img, n, logits, tag):
        """D(img) for n NHWC images; logits: fp32 view [16, ld] (row 0 receives the n logits).  Returns saved activations."""
        dc = self.dc
        sv = {"n": n, "img": img}
        x, hw, cin = img, 64, self.ch
        for i in range(4):
            hw //= 2
            col = self._buf(tag + "col%d" % i, n * hw * hw, 16 * cin)
            _im2col(x, n, 2 * hw, 2 * hw, cin, col)
            c = self._buf(tag + "c%d" % i, n * hw * hw, dc[i])
            sv["col%d" % i] = col
            if i == 0:
                gemm_bf16(col, self.D.bf["l1.weight"], c, "nt", act=3, act_slope=SLOPE)   # conv + LeakyReLU in the epilogue
                y = c
            else:
                gemm_bf16(col, self.D.bf["l%d.weight" % (i + 1)], c, "nt")
                y = self._buf(tag + "y%d" % i, c.shape[0], dc[i])
                st = self._buf(tag + "st%d" % i, 2, dc[i], torch.float32)
                _bn_fwd(c, self.D.view("bn%d.weight" % (i + 1)), self.D.view("bn%d.bias" % (i + 1)), ACT_LRELU, y, st, self.run_D[i])
                sv["c%d" % i], sv["st%d" % i] = c, st
            sv["y%d" % i] = y
            x, cin = y, dc[i]
        flat = x.view(n, 16 * dc[3])
        sv["flat"] = flat
        gemm_bf16(flat, self.D.bf["l5.weight"], logits, "nt", transpose=True)             # fp32 [16, ld]: row 0 = logits
        return sv

    def d_backward(self, sv, ds, grads, need_wgrad=True, need_dimg=False, tag="d"):
        """ds [n] fp32 = dL/dlogit.  Accumulates nothing: writes this pass's D gradient into `grads` (flat, D layout) when
        need_wgrad; returns dL/d(pre-sigmoid generator output) when need_dimg (img must then be a generator output)."""
        n, dc, D = sv["n"], self.dc, self.D
        dy5 = self._buf(tag + "dy5", n, 16)
        check(self.h, lib().gm_pack_col0(self.h, _ptr(ds), n, _ptr(dy5), 16, _stream()))
        if need_wgrad:
            gemm_bf16(dy5, sv["flat"], D.view("l5.weight", grads), "tn")                  # [16, 128h]
        dflat = self._buf(tag + "dflat", n, 16 * dc[3])
        gemm_bf16(dy5, D.bf_t["l5.weight"], dflat, "nt", K=16)
        d, hw = dflat.view(n * 16, dc[3]), 4
        for i in range(3, 0, -1):
            c, st, col = sv["c%d" % i], sv["st%d" % i], sv["col%d" % i]
            dcv = self._buf(tag + "dc%d" % i, c.shape[0], dc[i])
            dgb = self._buf(tag + "dgb%d" % i, 2, dc[i], torch.float32)
            _bn_bwd(d, c, st, D.view("bn%d.weight" % (i + 1)), D.view("bn%d.bias" % (i + 1)), ACT_LRELU, dcv, dgb)
            if need_wgrad:
                D.view("bn%d.bias" % (i + 1), grads).copy_(dgb[0])
                D.view("bn%d.weight" % (i + 1), grads).copy_(dgb[1])
                gemm_bf16(dcv, col, D.view("l%d.weight" % (i + 1), grads), "tn")          # [Cout, 16 Cin]
            dcol = self._buf(tag + "dcol%d" % i, col.shape[0], col.shape[1])
            gemm_bf16(dcv, D.bf_t["l%d.weight" % (i + 1)], dcol, "nt")
            dprev = self._buf(tag + "dprev%d" % i, n * 4 * hw * hw, dc[i - 1])
            # layer i's input is y_{i-1}: a BN layer's output (its backward applies LeakyReLU') or, for i == 1, lrelu(c_0)
            _col2im(dcol, n, hw, hw, dc[i - 1], dprev, C2I_LRELU_GRAD if i == 1 else C2I_NONE, sv["y0"] if i == 1 else None)
            d, hw = dprev, 2 * hw
        if need_wgrad:
            gemm_bf16(d, sv["col0"], D.view("l1.weight", grads), "tn")                    # [h, 16 ch]
        if not need_dimg:
            return None
        dcol = self._buf(tag + "dcol0", sv["col0"].shape[0], sv["col0"].shape[1])
        gemm_bf16(d, D.bf_t["l1.weight"], dcol, "nt")
        dpre = self._buf(tag + "dpre", n * 4096, self.ch)
        _col2im(dcol, n, 32, 32, self.ch, dpre, C2I_SIGMOID_GRAD, sv["img"])
        return dpre

    # ------------------------------------------------------------------ the train step (src/ns_gan.py:126-156)
    def stage_images(self, images):
        """[n, ch*64*64] flat (the reference's process_batch layout: NCHW flattened, src/ns_gan.py:225) or
        [n, ch, 64, 64] -> NHWC bf16 rows [n*4096, ch]."""
        n = images.shape[0]
        x = images.view(n, self.ch, 64, 64).permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        return x.view(n * 4096, self.ch)

    def d_grad(self, img_real, n, noise=None, inv_global_batch=None, seed=0, step=0):
        """train_D + backward (src/ns_gan.py:172-194,138): img_real NHWC rows of n images (stage_images).  Writes the flat
        D gradient (self.D.grads) and loss_buf[0]."""
        inv = 1.0 / n if inv_global_batch is None else inv_global_batch
        fake, gsv = self.g_forward(n, noise, seed, 2 * step)
        lr_, lf_ = self._buf("logits_r", 16, n, torch.float32), self._buf("logits_f", 16, n, torch.float32)
        sr = self.d_forward(img_real, n, lr_, "dr")
        sf = self.d_forward(fake, n, lf_, "df")
        logits = self._buf("logits", 1, 2 * n, torch.float32)[0]           # [real | fake], what the loss kernel walks
        logits[:n].copy_(lr_[0])
        logits[n:].copy_(lf_[0])
        ds = self._buf("ds", 1, 2 * n, torch.float32)[0]
        check(self.h, lib().gm_loss_rows(self.h, VARIANTS[self.variant], 0, _ptr(logits), n, 0, inv, _ptr(ds), None,
                                         _ptr(self.loss_buf), _stream()))
        g2 = self._buf("dgrad2", 1, self.D.total, torch.float32)[0]
        self.D.grads.zero_()
        g2.zero_()
        self.d_backward(sr, ds[:n], self.D.grads, tag="dr")
        self.d_backward(sf, ds[n:], g2, tag="df")
        self.D.grads.add_(g2)
        self.scores_ = logits
        return self.loss_buf[0]

    def g_grad(self, n, noise=None, inv_global_batch=None, seed=0, step=0):
        """train_G + backward (src/ns_gan.py:196-216,155): G gradients only."""
        inv = 1.0 / n if inv_global_batch is None else inv_global_batch
        fake, gsv = self.g_forward(n, noise, seed, 2 * step + 1)
        logits = self._buf("logits_g", 16, n, torch.float32)
        sf = self.d_forward(fake, n, logits, "df")
        ds = self._buf("ds_g", 1, n, torch.float32)[0]
        check(self.h, lib().gm_loss_rows(self.h, VARIANTS[self.variant], 0, _ptr(logits), n, 1, inv, _ptr(ds), None,
                                         C.c_void_p(self.loss_buf.data_ptr() + 4), _stream()))
        dpre = self.d_backward(sf, ds, None, need_wgrad=False, need_dimg=True, tag="df")
        self.g_backward(gsv, dpre)
        return self.loss_buf[1]

    def apply(self, net, hp):
        (self.G if net == 0 else self.D).adam(hp)

    def generate(self, noise):
        n = noise.shape[0]
        img, _ = self.g_forward(n, noise.float().contiguous(), tag="gen")
        return img.view(n, 64, 64, self.ch).permute(0, 3, 1, 2).float().reshape(n, -1)

    def discriminate(self, images):
        n = images.shape[0]
        logits = self._buf("logits_i", 16, n, torch.float32)
        self.d_forward(self.stage_images(images), n, logits, "di")
        return torch.sigmoid(logits[0, :n]).view(n, 1)
