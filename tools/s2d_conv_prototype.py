"""Design check for DESIGN.md §6b `Next`: a 4x4 stride-2 pad-1 convolution WITHOUT a column matrix and without TMA im2col
mode.  (CPU, float64, plain torch; not used by the product.)

Pad x [B, H, W, C] by one pixel and fold 2x2 pixel blocks into channels ("space to depth"):

    X'[b, h', w', (r, t, c)] = xpad[b, 2h'+r, 2w'+t, c]        h' < H' = H/2 + 1,  w' < W' = W/2 + 1

Then with kh = 2a + r, kw = 2b + t the convolution is a 2x2 stride-1 convolution of X', and on the flattened row index
m = (b H' + h') W' + w' its four taps are ROW OFFSETS of the same matrix:

    Y'[m, :] = sum_{a,b in {0,1}}  X'[m + a W' + b, :] @ Wt[a, b]          Wt[a, b][(r, t, c), co] = W[co, c, 2a+r, 2b+t]

i.e. one GEMM with K = 4 taps x 4C whose A-operand tile for k-block group (a, b) is the SAME 2-D tensor map at row
coordinate m0 + a W' + b (rows past the end: TMA zero fill).  Rows with h' = H/2 or w' = W/2 are garbage and are dropped
(or never stored) by the epilogue.  The gradients have the same form:

    dX'[m, :]   = sum_{a,b} dY'[m - a W' - b, :] @ Wt[a, b]^T     (dY' zero on the garbage rows and for negative rows)
    dWt[a, b]   = sum_m X'[m + a W' + b, :]^T dY'[m, :]

and the transposed convolution of the generator is the dX form with its input on the padded grid.  This script checks all
four against torch.nn.functional on random data (see tests/test_host_logic_cpu.py::test_space_to_depth_conv_identity)."""
import torch
import torch.nn.functional as F


def space_to_depth(x):
    """x [B, H, W, C] -> X' [B*(H/2+1)*(W/2+1), 4C] (rows = padded-grid pixels, channels (r, t, c))."""
    B, H, W, C = x.shape
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))                                     # [B, H+2, W+2, C]
    Hp, Wp = H // 2 + 1, W // 2 + 1
    return xp.view(B, Hp, 2, Wp, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B * Hp * Wp, 4 * C), Hp, Wp


def depth_to_space(Xs, B, H, W, C):
    """inverse of space_to_depth, dropping the one-pixel border: [B*(H/2+1)*(W/2+1), 4C] -> [B, H, W, C]."""
    Hp, Wp = H // 2 + 1, W // 2 + 1
    xp = Xs.view(B, Hp, Wp, 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * Hp, 2 * Wp, C)
    return xp[:, 1:H + 1, 1:W + 1]


def tap_weights(Wc):
    """conv weight [Co, C, 4, 4] -> Wt[a][b] = [(r, t, c), Co]."""
    Co, C = Wc.shape[:2]
    w = Wc.permute(2, 3, 1, 0).reshape(2, 2, 2, 2, C, Co)                 # [a, r, b, t, c, co]
    return [[w[a, :, b].reshape(4 * C, Co) for b in range(2)] for a in range(2)]


def shifted(M, off):
    """rows m -> M[m + off] with zero fill outside (what a TMA box at row coordinate m0 + off delivers)."""
    out = torch.zeros_like(M)
    n = M.shape[0]
    if off >= 0:
        out[:n - off] = M[off:]
    else:
        out[-off:] = M[:n + off]
    return out


def valid_rows(B, Hp, Wp, dtype):
    v = torch.ones(B, Hp, Wp, dtype=dtype)
    v[:, Hp - 1] = 0
    v[:, :, Wp - 1] = 0
    return v.view(-1, 1)


def conv_forward(x, Wc):
    """y [B, H/2, W/2, Co] as four row-shifted GEMMs."""
    B, H, W, _ = x.shape
    Xs, Hp, Wp = space_to_depth(x)
    Wt = tap_weights(Wc)
    Y = sum(shifted(Xs, a * Wp + b) @ Wt[a][b] for a in range(2) for b in range(2))
    return Y.view(B, Hp, Wp, -1)[:, :Hp - 1, :Wp - 1]


def conv_backward(x, Wc, dy):
    """-> (dx [B, H, W, C], dW [Co, C, 4, 4]) from dy [B, H/2, W/2, Co]."""
    B, H, W, C = x.shape
    Co = Wc.shape[0]
    Xs, Hp, Wp = space_to_depth(x)
    Wt = tap_weights(Wc)
    dYp = torch.zeros(B, Hp, Wp, Co, dtype=x.dtype)
    dYp[:, :Hp - 1, :Wp - 1] = dy                                           # zero on the garbage rows
    dYp = dYp.view(-1, Co)
    dXs = sum(shifted(dYp, -(a * Wp + b)) @ Wt[a][b].t() for a in range(2) for b in range(2))
    dWt = [[shifted(Xs, a * Wp + b).t() @ dYp for b in range(2)] for a in range(2)]
    dW = torch.stack([torch.stack([dWt[a][b].view(2, 2, C, Co) for b in range(2)]) for a in range(2)])   # [a, b, r, t, c, co]
    dW = dW.permute(5, 4, 0, 2, 1, 3).reshape(Co, C, 4, 4)
    return depth_to_space(dXs, B, H, W, C), dW


def conv_transpose_forward(x, Wm):
    """ConvTranspose2d(k4, s2, p1): x [B, Hi, Wi, Ci], weight [Ci, Co, 4, 4] -> [B, 2Hi, 2Wi, Co]: the dX form."""
    B, Hi, Wi, Ci = x.shape
    Co = Wm.shape[1]
    Hp, Wp = Hi + 1, Wi + 1
    Wt = tap_weights(Wm)                       # [Ci, Co, 4, 4] read as the weight of the adjoint convolution Co -> Ci
    Xp = torch.zeros(B, Hp, Wp, Ci, dtype=x.dtype)
    Xp[:, :Hi, :Wi] = x
    Xp = Xp.view(-1, Ci)
    Ys = sum(shifted(Xp, -(a * Wp + b)) @ Wt[a][b].t() for a in range(2) for b in range(2))
    return depth_to_space(Ys, B, 2 * Hi, 2 * Wi, Co)


def check(seed=0, B=2, H=8, W=12, C=3, Co=5):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, H, W, C, generator=g, dtype=torch.float64)
    Wc = torch.randn(Co, C, 4, 4, generator=g, dtype=torch.float64)
    xt = x.permute(0, 3, 1, 2).clone().requires_grad_()
    wt = Wc.clone().requires_grad_()
    yt = F.conv2d(xt, wt, stride=2, padding=1)
    err = {"conv": float((conv_forward(x, Wc) - yt.detach().permute(0, 2, 3, 1)).abs().max())}
    dy = torch.randn(B, H // 2, W // 2, Co, generator=g, dtype=torch.float64)
    yt.backward(dy.permute(0, 3, 1, 2))
    dx, dW = conv_backward(x, Wc, dy)
    err["dx"] = float((dx - xt.grad.permute(0, 2, 3, 1)).abs().max())
    err["dW"] = float((dW - wt.grad).abs().max())
    Wm = torch.randn(C, Co, 4, 4, generator=g, dtype=torch.float64)
    xi = torch.randn(B, H // 2, W // 2, C, generator=g, dtype=torch.float64)
    ref = F.conv_transpose2d(xi.permute(0, 3, 1, 2), Wm, stride=2, padding=1).permute(0, 2, 3, 1)
    err["convT"] = float((conv_transpose_forward(xi, Wm) - ref).abs().max())
    return err


if __name__ == "__main__":
    print(check())
