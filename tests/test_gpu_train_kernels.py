"""GPU: the backward kernels of DyGFormer's training path (csrc/train.cu) against float64 torch autograd of the same ops."""
import numpy as np
import pytest
import torch

from dyglib_b200 import ops
from dyglib_b200 import autograd as ag

pytestmark = pytest.mark.gpu


def rel(got, want):
    return float((got.double() - want).abs().max() / want.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (37, 5, 9), (1000, 200, 800), (25600, 800, 200), (4099, 172, 616), (300, 1, 172)])
def test_gemm_dw_vs_float64(M, N, K):
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    G = torch.randn(M, N, device='cuda', generator=g)
    X = torch.randn(M, K, device='cuda', generator=g)
    dw, db = ops.gemm_dw(G, X, want_bias=True)
    tol = 2e-5 if M < 512 else 2e-4        # from 512 rows on: BF16x3 on mma.sync (dropped terms ~2^-16 per product over M products)
    assert rel(dw, G.double().t() @ X.double()) < tol
    assert rel(db, G.double().sum(0)) < 2e-5
    if M >= 1000:
        # the tcgen05 route for very large layers (planes of G^T and X^T): BF16x3 over a contraction of M terms
        old, ops.GEMM_DW_TC_FLOPS = ops.GEMM_DW_TC_FLOPS, 0.0
        try:
            dw2, db2 = ops.gemm_dw(G, X, want_bias=True)
        finally:
            ops.GEMM_DW_TC_FLOPS = old
        assert rel(dw2, G.double().t() @ X.double()) < 3e-4 and rel(db2, G.double().sum(0)) < 2e-5
    # accumulation into a column window of a wider gradient (concatenated inputs of one layer)
    wide = torch.zeros(N, K + 7, device='cuda')
    ops.gemm_dw(G, X, dw=wide[:, 3:3 + K])
    assert rel(wide[:, 3:3 + K], G.double().t() @ X.double()) < tol and float(wide[:, :3].abs().max()) == 0.0


@pytest.mark.parametrize('M,D', [(1, 8), (77, 200), (5000, 172), (333, 272), (64, 512)])
def test_layernorm_bwd_vs_float64(M, D):
    g = torch.Generator(device='cuda').manual_seed(M + D)
    x = (torch.randn(M, D, device='cuda', generator=g) * 2 + 0.5)
    gamma = torch.randn(D, device='cuda', generator=g) * 0.3 + 1
    beta = torch.randn(D, device='cuda', generator=g) * 0.3
    dy = torch.randn(M, D, device='cuda', generator=g)
    xd, gd, bd = (t.double().requires_grad_(True) for t in (x, gamma, beta))
    torch.nn.functional.layer_norm(xd, (D,), gd, bd, 1e-5).backward(dy.double())
    dx, dg, db = ops.layernorm_bwd(x, gamma, 1e-5, dy)
    assert rel(dx, xd.grad) < 2e-5 and rel(dg, gd.grad) < 2e-5 and rel(db, bd.grad) < 2e-5
    # the autograd wrapper: forward equals torch, gradients flow
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y = ag.layer_norm(xr, gr, br, 1e-5)
    assert rel(y.detach(), torch.nn.functional.layer_norm(x.double(), (D,), gamma.double(), beta.double(), 1e-5)) < 1e-5
    y.backward(dy)
    assert rel(xr.grad, xd.grad) < 2e-5 and rel(gr.grad, gd.grad) < 2e-5


def test_gelu_fwd_bwd_vs_float64():
    g = torch.Generator(device='cuda').manual_seed(3)
    v = torch.randn(301, 800, device='cuda', generator=g) * 3
    mask = (torch.rand(301, 800, device='cuda', generator=g) > 0.2).float() / 0.8
    dh = torch.randn(301, 800, device='cuda', generator=g)
    vd = v.double().requires_grad_(True)
    (torch.nn.functional.gelu(vd) * mask.double()).backward(dh.double())
    h, hs = ops.gelu_fwd(v, mask, want='both')
    want = torch.nn.functional.gelu(v.double()) * mask.double()
    assert rel(h, want) < 1e-6 and rel(hs.float(), want) < 3e-5
    assert rel(ops.gelu_bwd(v, mask, dh), vd.grad) < 2e-6
    assert rel(ops.gelu_bwd(v, None, dh), torch.autograd.grad(torch.nn.functional.gelu(vd).sum(), vd)[0] * dh.double()) < 2e-6


@pytest.mark.parametrize('B,S,H,hd,drop', [(3, 64, 2, 100, False), (5, 18, 2, 100, True), (2, 33, 4, 24, True), (4, 1, 2, 8, False)])
def test_seq_attention_train_vs_float64(B, S, H, hd, drop):
    g = torch.Generator(device='cuda').manual_seed(B * 100 + S)
    D = H * hd
    qkv = torch.randn(B * S, 3 * D, device='cuda', generator=g)
    mask = ((torch.rand(B, H, S, S, device='cuda', generator=g) > 0.3).float() / 0.7) if drop else None
    do = torch.randn(B * S, D, device='cuda', generator=g)
    qd = qkv.double().requires_grad_(True)
    q, k, v = (qd.reshape(B, S, 3, H, hd)[:, :, i].transpose(1, 2) for i in range(3))
    p = torch.softmax((q @ k.transpose(-1, -2)) / np.sqrt(hd), dim=-1)
    pm = p * mask.double() if drop else p
    want = (pm @ v).transpose(1, 2).reshape(B * S, D)
    want.backward(do.double())
    out, probs = ops.seq_attention_train_fwd(qkv, B, S, H, hd, mask)
    assert rel(out, want.detach()) < 1e-5 and rel(probs, p.detach()) < 1e-5
    assert rel(ops.seq_attention_train_bwd(qkv, B, S, H, hd, mask, probs, do), qd.grad) < 2e-5


def test_linear_backward_runs_on_own_kernels():
    g = torch.Generator(device='cuda').manual_seed(11)
    x1 = torch.randn(500, 172, device='cuda', generator=g, requires_grad=True)
    x2 = torch.randn(500, 100, device='cuda', generator=g, requires_grad=True)
    w = torch.randn(64, 272, device='cuda', generator=g, requires_grad=True)
    b = torch.randn(64, device='cuda', generator=g, requires_grad=True)
    gy = torch.randn(500, 64, device='cuda', generator=g)
    ag.linear([x1, x2], w, b, act=ops.ACT_RELU).backward(gy)
    xd = torch.cat([x1, x2], 1).detach().double().requires_grad_(True)
    wd, bd = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    torch.relu(xd @ wd.t() + bd).backward(gy.double())
    assert rel(torch.cat([x1.grad, x2.grad], 1), xd.grad) < 5e-5
    assert rel(w.grad, wd.grad) < 5e-5 and rel(b.grad, bd.grad) < 5e-5


@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (37, 5, 9), (700, 272, 444), (25600, 800, 200)])
def test_gemm_dx_small_and_tcgen05_paths(M, N, K):
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    G = torch.randn(M, N, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g)
    assert rel(ops.gemm_dx(G, W), G.double() @ W.double()) < 3e-5
    Y = torch.randn(M, N, device='cuda', generator=g)
    assert rel(ops.gemm_dx(G, W, Y), (G * (Y > 0)).double() @ W.double()) < 3e-5


@pytest.mark.parametrize('M,N,K1,K2,relu', [(1, 1, 1, 0, False), (500, 64, 172, 100, True), (803, 172, 172, 272, False), (77, 1, 172, 0, True)])
def test_linear_bwd_one_launch(M, N, K1, K2, relu):
    """dyg_linear_bwd: dX, dW, db of a small layer (with concatenated inputs and the ReLU mask) against float64."""
    g = torch.Generator(device='cuda').manual_seed(M + N + K1)
    widths = [K1] + ([K2] if K2 else [])
    xs = [torch.randn(M, k, device='cuda', generator=g) for k in widths]
    w = torch.randn(N, sum(widths), device='cuda', generator=g)
    gy = torch.randn(M, N, device='cuda', generator=g)
    y = torch.randn(M, N, device='cuda', generator=g) if relu else None
    dx, dw, db = ops.linear_bwd(gy, y, xs, widths, w, True, True, True)
    gm = (gy * (y > 0)).double() if relu else gy.double()
    assert rel(dx, gm @ w.double()) < 2e-5
    assert rel(dw, gm.t() @ torch.cat(xs, 1).double()) < 2e-5
    assert rel(db, gm.sum(0)) < 2e-5
    dx2, dw2, db2 = ops.linear_bwd(gy, y, xs, widths, w, True, False, False)
    assert dw2 is None and db2 is None and rel(dx2, gm @ w.double()) < 2e-5
