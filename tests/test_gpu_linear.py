"""GPU: dyg_linear (FFMA) and dyg_linear_tc (tcgen05 BF16x3) against a float64 contraction of the same inputs."""
import numpy as np
import pytest
import torch

from dyglib_b200 import ops

pytestmark = pytest.mark.gpu


def rel_err(got, want):
    return float((got.double() - want).abs().max() / want.abs().max().clamp_min(1e-30))


SHAPES = [(128, 16, 64), (129, 50, 172), (1000, 200, 200), (4096, 600, 200), (777, 800, 200), (640, 200, 800),
          (2000, 272, 888), (513, 888, 172), (300, 516, 616), (260, 172, 444), (5000, 50, 344), (256, 1, 172)]


@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('tc', [False, True])
def test_dense_linear(M, N, K, tc):
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    a = torch.randn(M, K, device='cuda', generator=g)
    w = torch.randn(N, K, device='cuda', generator=g) / np.sqrt(K)
    b = torch.randn(N, device='cuda', generator=g)
    r = torch.randn(M, N, device='cuda', generator=g)
    want = a.double() @ w.double().t() + b.double() + r.double()
    got = ops.linear([ops.seg_rows(a)], M, w, b, residual=r, tc=tc)
    e = rel_err(got, want)
    assert e < (2e-4 if tc else 2e-5), e
    for act, fn in ((ops.ACT_RELU, torch.relu), (ops.ACT_GELU, torch.nn.functional.gelu), (ops.ACT_SIGMOID, torch.sigmoid)):
        got = ops.linear([ops.seg_rows(a)], M, w, b, act=act, tc=tc)
        want_a = fn(a.double() @ w.double().t() + b.double())
        assert rel_err(got, want_a) < (2e-4 if tc else 2e-5)


@pytest.mark.parametrize('tc', [False, True])
def test_segmented_gathered_linear(tc):
    """[patch-gathered rows (two tables added) | time encoding with mask | dense] with output row remapping."""
    g = torch.Generator(device='cuda').manual_seed(7)
    P, F, T, Dd = 2, 172, 100, 52
    B, Lp = 37, 16
    ntok = Lp // P
    M = B * ntok
    tab = torch.randn(500, F, device='cuda', generator=g)
    tab2 = torch.randn(500, F, device='cuda', generator=g)
    idx = torch.randint(0, 500, (B * Lp,), device='cuda', generator=g)
    idx[::7] = 0
    tn = (torch.rand(B * Lp, device='cuda', generator=g) * 1e5).float()
    tq = (1e5 + torch.rand(B, device='cuda', generator=g, dtype=torch.float64) * 1e5)
    w_t = (1.0 / 10 ** torch.linspace(0, 9, T, device='cuda')).float().contiguous()
    b_t = (0.1 * torch.randn(T, device='cuda', generator=g)).contiguous()
    dense = torch.randn(M, Dd, device='cuda', generator=g)
    K = P * F + P * T + Dd
    N = 50
    W = torch.randn(N, K, device='cuda', generator=g) / np.sqrt(K)
    bias = torch.randn(N, device='cuda', generator=g)
    S = ntok + 5
    out = torch.zeros(B * S, 200, device='cuda')
    segs = [ops.seg_rows(tab, F, idx, group=P, table2=tab2),
            ops.seg_time(tn, w_t, b_t, mask_ids=idx, group=P, t_query=tq, tq_div=Lp),
            ops.seg_rows(dense)]
    ops.linear(segs, M, W, bias, out=out[:, 50:100], c_group=ntok, c_group_stride=S, c_offset=3, tc=tc)
    rows = (tab[idx] + tab2[idx]).double().reshape(M, P * F)
    dt = (tq.repeat_interleave(Lp) - tn.double()).float()
    te = torch.cos(torch.addcmul(b_t.double()[None, :], dt.double()[:, None], w_t.double()[None, :]).float().double())
    # the fp32 FMA argument is what the kernel uses; recompute it exactly in fp32
    arg = torch.tensor(np.float32(dt.cpu().numpy()[:, None]) * np.float32(w_t.cpu().numpy()[None, :]), device='cuda')  # not fused
    te = torch.cos((dt[:, None].double() * w_t[None, :].double() + b_t[None, :].double()).float().double())
    te = te * (idx != 0).double()[:, None]
    A = torch.cat([rows, te.reshape(M, P * T), dense.double()], dim=1)
    want = A @ W.double().t() + bias.double()
    got = out.reshape(B, S, 200)[:, 3:3 + ntok, 50:100].reshape(M, N)
    # time-encoding arguments reach 2e5 rad: allow for the fp32 argument rounding of single FMA vs float64 here
    assert rel_err(got, want) < 5e-3
    # untouched parts of the output stay zero
    assert float(out[:, :50].abs().max()) == 0 and float(out[:, 100:].abs().max()) == 0
    assert float(out.reshape(B, S, 200)[:, :3].abs().max()) == 0
    # FFMA and tensor-core paths agree tightly with each other
    out2 = torch.zeros_like(out)
    ops.linear(segs, M, W, bias, out=out2[:, 50:100], c_group=ntok, c_group_stride=S, c_offset=3, tc=not tc)
    assert rel_err(out, out2.double()) < 2e-4
