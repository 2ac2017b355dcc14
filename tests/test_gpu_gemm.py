"""GPU: dyg_gemm_bf16x3 (TMA + tcgen05, BF16x3 operand planes), dyg_split_bf16 and dyg_layernorm_split against
float64 torch references of the same inputs."""
import numpy as np
import pytest
import torch

from dyglib_b200 import ops

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    return float((got.double() - want).abs().max() / want.abs().max().clamp_min(1e-30))


# (M, N, K): DyGFormer's transformer / output shapes, ragged M, N and K tails, single tiles, many tiles per CTA
SHAPES = [(128, 16, 32), (1, 8, 8), (129, 50, 172), (1000, 200, 200), (4096, 600, 200), (777, 800, 200), (640, 200, 800),
          (513, 172, 200), (300, 516, 616), (260, 272, 444), (40000, 200, 200), (25600, 800, 200), (256, 1, 176),
          (148 * 128 * 2 + 5, 208, 72)]


def test_split_roundtrip():
    g = torch.Generator(device='cuda').manual_seed(1)
    x = torch.randn(333, 200, device='cuda', generator=g) * 37.0
    s = ops.split_bf16(x)
    assert s.planes.shape == (2, 333, 208) and s.planes.dtype == torch.bfloat16
    hi = x.to(torch.bfloat16)
    mid = (x - hi.float()).to(torch.bfloat16)
    assert torch.equal(s.hi[:, :200], hi) and torch.equal(s.mid[:, :200], mid)
    assert rel_err(s.float(), x.double()) < 2e-5
    y = torch.randn(7, 5, device='cuda', generator=g)          # odd width, strided input
    big = torch.zeros(7, 9, device='cuda')
    big[:, :5] = y
    s = ops.split_bf16(big[:, :5])
    assert rel_err(s.float(), y.double()) < 2e-5


@pytest.mark.parametrize('M,N,K', SHAPES)
def test_gemm_vs_float64(M, N, K):
    g = torch.Generator(device='cuda').manual_seed(M + 3 * N + 7 * K)
    a = torch.randn(M, K, device='cuda', generator=g)
    w = torch.randn(N, K, device='cuda', generator=g) / np.sqrt(K)
    b = torch.randn(N, device='cuda', generator=g)
    r = torch.randn(M, N, device='cuda', generator=g)
    sa = ops.split_bf16(a)
    base = a.double() @ w.double().t()
    got = ops.gemm(sa, w)
    assert rel_err(got, base) < 3e-5, rel_err(got, base)
    want = base + b.double() + r.double()
    got, gs = ops.gemm(sa, w, b, residual=r, want='both')
    assert rel_err(got, want) < 3e-5
    assert rel_err(gs.float(), want) < 5e-5
    hi = got.to(torch.bfloat16)
    assert torch.equal(gs.hi[:, :N], hi) and torch.equal(gs.mid[:, :N], (got - hi.float()).to(torch.bfloat16))
    for act, fn in ((ops.ACT_RELU, torch.relu), (ops.ACT_GELU, torch.nn.functional.gelu), (ops.ACT_SIGMOID, torch.sigmoid)):
        gs = ops.gemm(sa, w, b, act=act, want='split')
        assert rel_err(gs.float(), fn(base + b.double())) < 5e-5


def test_gemm_chain_keeps_fp32_accuracy():
    """FFN shape chain 200 -> 800 (GELU, planes only) -> 200 (+ residual): errors stay at the 1e-5 level."""
    g = torch.Generator(device='cuda').manual_seed(5)
    M = 6400
    x = torch.randn(M, 200, device='cuda', generator=g)
    w0 = torch.randn(800, 200, device='cuda', generator=g) / np.sqrt(200)
    b0 = torch.randn(800, device='cuda', generator=g)
    w1 = torch.randn(200, 800, device='cuda', generator=g) / np.sqrt(800)
    b1 = torch.randn(200, device='cuda', generator=g)
    h = ops.gemm(ops.split_bf16(x), w0, b0, act=ops.ACT_GELU, want='split')
    out = ops.gemm(h, w1, b1, residual=x)
    want = torch.nn.functional.gelu(x.double() @ w0.double().t() + b0.double()) @ w1.double().t() + b1.double() + x.double()
    assert rel_err(out, want) < 5e-5


def test_gemm_row_independent_of_batch():
    """A row's result does not depend on how many rows share the launch (grouped == per-batch results)."""
    g = torch.Generator(device='cuda').manual_seed(9)
    a = torch.randn(1000, 200, device='cuda', generator=g)
    w = torch.randn(600, 200, device='cuda', generator=g)
    full = ops.gemm(ops.split_bf16(a), w)
    part = ops.gemm(ops.split_bf16(a[300:437].contiguous()), w)
    assert torch.equal(full[300:437], part)


@pytest.mark.parametrize('M,D', [(1, 200), (1000, 200), (77, 64), (300, 444), (50, 800)])
def test_layernorm_split(M, D):
    g = torch.Generator(device='cuda').manual_seed(M + D)
    x = torch.randn(M, D, device='cuda', generator=g) * 3 + 1
    gm = torch.randn(D, device='cuda', generator=g)
    bt = torch.randn(D, device='cuda', generator=g)
    y = torch.empty_like(x)
    s = ops.layernorm_split(x, gm, bt, eps=1e-5, y=y)
    want = torch.nn.functional.layer_norm(x.double(), (D,), gm.double(), bt.double(), 1e-5)
    assert rel_err(y, want) < 1e-5
    assert rel_err(s.float(), want) < 3e-5


@pytest.mark.parametrize('B,S,H,hd', [(3, 64, 2, 100), (5, 20, 2, 100), (2, 7, 1, 32), (4, 100, 2, 64), (2, 128, 2, 128),
                                       (300, 64, 2, 100), (3, 33, 2, 50), (2, 2, 2, 100)])
def test_seq_attention_tc(B, S, H, hd):
    g = torch.Generator(device='cuda').manual_seed(B + S + hd)
    D = H * hd
    qkv = torch.randn(B * S, 3 * D, device='cuda', generator=g) * 1.5
    out, sp = ops.seq_attention_tc(qkv, B, S, H, hd, want='both')
    q, k, v = (qkv.double()[:, i * D:(i + 1) * D].reshape(B, S, H, hd).transpose(1, 2) for i in range(3))
    a = torch.softmax(q @ k.transpose(-1, -2) / np.sqrt(hd), dim=-1)
    want = (a @ v).transpose(1, 2).reshape(B * S, D)
    assert rel_err(out, want) < 3e-5, rel_err(out, want)
    assert rel_err(sp.float(), want) < 5e-5
    old = ops.seq_attention(qkv, B, S, H, hd)
    assert rel_err(old, want) < 1e-5


@pytest.mark.parametrize('P,B,ns,nd', [(2, 37, 8, 5), (1, 9, 32, 32), (16, 11, 3, 2), (2, 200, 32, 32), (4, 3, 1, 1)])
def test_patch_project_vs_float64(P, B, ns, nd):
    """Fused gather + time encoding + patching + four channel projections against the explicit float64 formula."""
    g = torch.Generator(device='cuda').manual_seed(100 * P + B)
    F, T, C = 172, 100, 50
    N, E, maxc = 400, 900, 40
    node = torch.randn(N, F, device='cuda', generator=g)
    edge = torch.randn(E, F, device='cuda', generator=g)
    node[0] = 0
    edge[0] = 0
    lut = torch.randn(maxc + 1, C, device='cuda', generator=g)
    tw = (1.0 / 10 ** torch.linspace(0, 9, T, device='cuda')).float().contiguous()
    tb = (0.1 * torch.randn(T, device='cuda', generator=g)).contiguous()
    ws = [torch.randn(C, P * w, device='cuda', generator=g) / np.sqrt(P * w) for w in (F, F, T, C)]
    bias = torch.randn(4 * C, device='cuda', generator=g)
    tq = 2e5 + torch.rand(B, device='cuda', generator=g, dtype=torch.float64) * 1e5
    S = ns + nd
    X = torch.full((B * S, 4 * C), float('nan'), device='cuda')
    sides, want = [], torch.zeros(B, S, 4 * C, dtype=torch.float64, device='cuda')
    for ntok, off in ((ns, 0), (nd, ns)):
        Lp = ntok * P
        ids = torch.randint(0, N, (B, Lp), device='cuda', generator=g)
        ids[:, Lp // 2:][torch.rand(B, Lp - Lp // 2, device='cuda', generator=g) < 0.5] = 0
        eids = torch.randint(0, E, (B, Lp), device='cuda', generator=g)
        tn = (torch.rand(B, Lp, device='cuda', generator=g) * 2e5).float()
        ca = torch.randint(0, maxc + 1, (B, Lp), device='cuda', generator=g)
        cb = torch.randint(0, maxc + 1, (B, Lp), device='cuda', generator=g)
        sides.append((ids, eids, tn, ca, cb, ntok, off))
        dt = (tq[:, None] - tn.double()).float()
        te = torch.cos(torch.addcmul(tb[None, None, :], dt[:, :, None], tw[None, None, :]).double())   # fp32 fma argument
        te = te * (ids != 0).double()[:, :, None]
        feats = [node[ids].double(), edge[eids].double(), te, (lut[ca] + lut[cb]).double()]
        for ch, (f, w) in enumerate(zip(feats, ws)):
            patches = f.reshape(B, ntok, P * f.shape[-1])
            want[:, off:off + ntok, ch * C:(ch + 1) * C] = patches @ w.double().t() + bias[ch * C:(ch + 1) * C].double()
    packed = ops.pack_patch_weights(ws[0], ws[1], ws[2], ws[3], P)
    ops.patch_project(sides, ops.table_planes(node), F, ops.table_planes(edge), F, ops.table_planes(lut), C, tq, tw, tb, packed,
                      bias, P, C, S, X, zero_rows=3 if P != 4 else 0)
    got = X.reshape(B, S, 4 * C)
    assert not torch.isnan(got).any()
    # node / edge / co-occurrence channels are pure BF16x3 contractions
    for ch in (0, 1, 3):
        assert rel_err(got[..., ch * C:(ch + 1) * C], want[..., ch * C:(ch + 1) * C]) < 5e-5, ch
    # time channel: torch's fp32 addcmul may round the argument differently from a single FMA (arguments reach 3e5 rad)
    assert rel_err(got[..., 2 * C:3 * C], want[..., 2 * C:3 * C]) < 5e-3


@pytest.mark.parametrize('M,D,Dff', [(256, 200, 800), (1, 200, 800), (300, 200, 800), (5000, 200, 800), (777, 64, 32), (1000, 208, 256), (333, 8, 96),
                                     (148 * 256 + 77, 200, 800)])
def test_ln_ffn_fused(M, D, Dff):
    g = torch.Generator(device='cuda').manual_seed(M + D + Dff)
    x = torch.randn(M, D, device='cuda', generator=g) * 2 + 0.5
    gm = 1 + 0.1 * torch.randn(D, device='cuda', generator=g)
    bt = 0.1 * torch.randn(D, device='cuda', generator=g)
    w1 = torch.randn(Dff, D, device='cuda', generator=g) / np.sqrt(D)
    b1 = torch.randn(Dff, device='cuda', generator=g)
    w2 = torch.randn(D, Dff, device='cuda', generator=g) / np.sqrt(Dff)
    b2 = torch.randn(D, device='cuda', generator=g)
    got = ops.ln_ffn(x, gm, bt, 1e-5, w1, b1, w2, b2)
    xd = x.double()
    y = torch.nn.functional.layer_norm(xd, (D,), gm.double(), bt.double(), 1e-5)
    want = xd + torch.nn.functional.gelu(y @ w1.double().t() + b1.double()) @ w2.double().t() + b2.double()
    assert rel_err(got, want) < 5e-5, rel_err(got, want)


def test_empty_and_tiny_inputs():
    """M = 0 is a no-op for every tensor-core entry point; single rows work (tiles are padded by TMA zero fill)."""
    dev = 'cuda'
    w = torch.randn(200, 200, device=dev)
    out = ops.gemm(ops.split_bf16(torch.zeros(0, 200, device=dev)), w)
    assert out.shape == (0, 200)
    x0 = torch.zeros(0, 200, device=dev)
    g1, b1 = torch.ones(200, device=dev), torch.zeros(200, device=dev)
    assert ops.ln_ffn(x0, g1, b1, 1e-5, torch.randn(800, 200, device=dev), torch.zeros(800, device=dev), torch.randn(200, 800, device=dev),
                      torch.zeros(200, device=dev)).shape == (0, 200)
    assert ops.layernorm_split(x0, g1, b1).rows == 0
    sp = ops.seq_attention_tc(torch.zeros(0, 600, device=dev), 0, 64, 2, 100, want='split')
    assert sp.rows == 0
    x1 = torch.randn(1, 200, device=dev)
    got = ops.gemm(ops.split_bf16(x1), w)
    assert rel_err(got, x1.double() @ w.double().t()) < 3e-5


def test_gemm_rejects_bad_arguments():
    a = ops.split_bf16(torch.randn(4, 200, device='cuda'))
    with pytest.raises(ValueError):
        ops.gemm(a, torch.randn(8, 100, device='cuda'))          # K mismatch
    with pytest.raises(ValueError):
        ops.ln_ffn(torch.randn(4, 50, device='cuda'), torch.ones(50, device='cuda'), torch.zeros(50, device='cuda'), 1e-5,
                   torch.randn(64, 50, device='cuda'), torch.zeros(64, device='cuda'), torch.randn(50, 64, device='cuda'),
                   torch.zeros(50, device='cuda'))                # D % 8 != 0


@pytest.mark.parametrize('M,N,D', [(1, 8, 8), (129, 50, 176), (1000, 816, 200), (40000, 816, 200), (777, 600, 200), (513, 224, 224),
                                   (148 * 128 * 2 + 5, 208, 72)])
def test_ln_gemm_vs_float64(M, N, D):
    """dyg_ln_gemm_bf16x3: LayerNorm fused into the projection GEMM, against float64 LayerNorm + matmul."""
    g = torch.Generator(device='cuda').manual_seed(M + 3 * N + 7 * D)
    x = torch.randn(M, D, device='cuda', generator=g) * 2.0 + 0.3
    w = torch.randn(N, D, device='cuda', generator=g) / np.sqrt(D)
    b = torch.randn(N, device='cuda', generator=g)
    gamma = torch.randn(D, device='cuda', generator=g) * 0.2 + 1.0
    beta = torch.randn(D, device='cuda', generator=g) * 0.2
    y = torch.nn.functional.layer_norm(x.double(), (D,), gamma.double(), beta.double(), 1e-5)
    want = y @ w.double().t() + b.double()
    got, gs = ops.ln_gemm(x, gamma, beta, 1e-5, w, b, want='both')
    assert rel_err(got, want) < 3e-5, rel_err(got, want)
    assert rel_err(gs.float(), want) < 5e-5
    # planes only, no activation: when N % 16 == 0 they leave through shared memory + TMA stores (epilogue_planes_tma_tile)
    gp = ops.ln_gemm(x, gamma, beta, 1e-5, w, b, want='split')
    assert rel_err(gp.float(), want) < 5e-5
    assert torch.equal(gp.hi[:, :N], gs.hi[:, :N]) and torch.equal(gp.mid[:, :N], gs.mid[:, :N])     # same bits as the direct-store epilogue
    gs = ops.ln_gemm(x, gamma, beta, 1e-5, w, b, act=ops.ACT_GELU, want='split')
    assert rel_err(gs.float(), torch.nn.functional.gelu(want)) < 5e-5
