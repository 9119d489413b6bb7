"""tcgen05 GEMM kernel vs a plain PyTorch fp32 reference of the same op (GPU).
Covers both operand forms (K-major "nt", MN-major "tn"), every fused epilogue,
ragged tiles (M, N, K not multiples of the tile), padding columns and split-K."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _mk(rows, cols, ld=None, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ld = ld or cols
    t = torch.zeros(rows, ld, device="cuda", dtype=torch.bfloat16)
    t[:, :cols] = (torch.randn(rows, cols, device="cuda", generator=g) * scale).to(torch.bfloat16)
    return t


@pytest.mark.parametrize("M,N,K", [(128, 208, 64), (128, 64, 64), (256, 400, 784), (192, 784, 400),
                                   (64, 400, 32), (1000, 416, 128), (128, 48, 400)])
def test_nt_plain(M, N, K):
    import gm_b200
    A, B = _mk(M, K, seed=1), _mk(N, K, seed=2)
    out = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    gm_b200.gemm_bf16(A, B, out, "nt")
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    err = _rel(out.float(), ref)
    assert err < 4e-3, (M, N, K, err)


def test_nt_epilogues():
    import gm_b200
    M, N, K, ldo = 320, 400, 784, 416
    A, B = _mk(M, K, 800, seed=3, scale=0.5), _mk(N, K, seed=4, scale=0.05)
    bias = torch.randn(N, device="cuda") * 0.1
    w = torch.randn(N, device="cuda")
    ref_pre = A[:, :K].float() @ B.float().t() + bias
    # relu + ones column + fused row-dot
    out = torch.full((M, ldo), 7.0, device="cuda", dtype=torch.bfloat16)
    slots = torch.zeros(4, M, device="cuda")
    gm_b200.gemm_bf16(A, B, out, "nt", K=K, bias=bias, act=1, pad_one=True, out_cols=ldo, dot_w=w, dot_out=slots)
    torch.cuda.synchronize()
    ref = torch.relu(ref_pre)
    assert _rel(out[:, :N].float(), ref) < 4e-3
    assert torch.all(out[:, N] == 1) and torch.all(out[:, N + 1:] == 0)
    assert _rel(slots.sum(0), ref @ w) < 2e-3
    # sigmoid
    out2 = torch.zeros(M, ldo, device="cuda", dtype=torch.bfloat16)
    gm_b200.gemm_bf16(A, B, out2, "nt", K=K, bias=bias, act=2)
    torch.cuda.synchronize()
    assert _rel(out2[:, :N].float(), torch.sigmoid(ref_pre)) < 4e-3
    assert torch.all(out2[:, N:] == 0)
    # aux modes
    aux = torch.rand(M, ldo, device="cuda").to(torch.bfloat16)
    out3 = torch.zeros(M, ldo, device="cuda", dtype=torch.bfloat16)
    gm_b200.gemm_bf16(A, B, out3, "nt", K=K, aux=aux, aux_mode=1)
    torch.cuda.synchronize()
    a = aux[:, :N].float()
    assert _rel(out3[:, :N].float(), (ref_pre - bias) * a * (1 - a)) < 4e-3
    auxm = (torch.randn(M, ldo, device="cuda")).to(torch.bfloat16)
    out4 = torch.zeros(M, ldo, device="cuda", dtype=torch.bfloat16)
    gm_b200.gemm_bf16(A, B, out4, "nt", K=K, aux=auxm, aux_mode=2)
    torch.cuda.synchronize()
    assert _rel(out4[:, :N].float(), (ref_pre - bias) * (auxm[:, :N].float() > 0)) < 4e-3


@pytest.mark.parametrize("K,M,N,tr", [(64, 128, 64, False), (64, 785, 400, True), (128, 784, 401, False),
                                      (64, 400, 21, False), (8192, 785, 400, True), (4096, 400, 21, False),
                                      (2048, 784, 401, False)])
def test_tn_splitk(K, M, N, tr):
    import gm_b200
    lda, ldb = ((M + 15) // 16) * 16, ((N + 15) // 16) * 16
    A, B = _mk(K, M, lda, seed=5, scale=0.3), _mk(K, N, ldb, seed=6, scale=0.3)
    ldc = ((max(M, N) + 63) // 64) * 64
    out = torch.full((N if tr else M, ldc), 7.0, device="cuda", dtype=torch.float32)
    gm_b200.gemm_bf16(A, B, out, "tn", M=M, N=N, transpose=tr)
    torch.cuda.synchronize()
    ref = A[:, :M].float().t() @ B[:, :N].float()
    got = out[:N, :M].t() if tr else out[:M, :N]
    err = _rel(got, ref)
    assert err < 1e-4, (K, M, N, tr, err)


def test_adam_step_matches_torch():
    import gm_b200
    n = 10007
    p = torch.randn(n, device="cuda")
    ref_p = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref_p], lr=2e-4)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    hp = gm_b200.AdamHP.make(2e-4)
    for step in range(1, 4):
        g = torch.randn(n, device="cuda") * 0.01
        ref_p.grad = g.clone()
        opt.step()
        gm_b200.adam_step(p, g, m, v, hp, step)
    torch.cuda.synchronize()
    np.testing.assert_allclose(p.cpu().numpy(), ref_p.detach().cpu().numpy(), rtol=1e-6, atol=1e-7)
