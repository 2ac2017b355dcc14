"""GPU: dyg_seq_attention_fold (tcgen05 attention with the out-projection folded into the value projection) against a
float64 torch nn.MultiheadAttention + residual of the same inputs (models/DyGFormer.py:442-455)."""
import pytest
import torch

from dyglib_b200 import ops

pytestmark = pytest.mark.gpu


def reference_block(x, B, S, mha, ln):
    xd = x.double().cpu().reshape(B, S, -1)
    m = torch.nn.MultiheadAttention(mha.embed_dim, mha.num_heads).double()
    m.load_state_dict({k: v.double().cpu() for k, v in mha.state_dict().items()})
    l = torch.nn.LayerNorm(mha.embed_dim).double()
    l.load_state_dict({k: v.double().cpu() for k, v in ln.state_dict().items()})
    with torch.no_grad():
        y = l(xd).transpose(0, 1)
        return (xd + m(y, y, y)[0].transpose(0, 1)).reshape(B * S, -1)


# (B, S, D, H): DyGFormer's block (D=200, H=2) with full / ragged / tiny sequences, one slot and several tiles per CTA
CASES = [(2, 64, 200, 2), (5, 64, 200, 2), (3, 50, 200, 2), (7, 32, 200, 2), (9, 18, 200, 2), (4, 2, 200, 2), (700, 64, 200, 2),
         (333, 40, 200, 2), (6, 64, 64, 2), (6, 33, 96, 4), (5, 64, 208, 2)]


@pytest.mark.parametrize('B,S,D,H', CASES)
def test_seq_attention_fold_vs_float64(B, S, D, H):
    torch.manual_seed(B * 1000 + S * 10 + H)
    mha = torch.nn.MultiheadAttention(D, H).cuda()
    ln = torch.nn.LayerNorm(D).cuda()
    with torch.no_grad():
        mha.in_proj_bias.normal_(0, 0.3)
        mha.out_proj.bias.normal_(0, 0.3)
        mha.in_proj_weight.mul_(2.0)           # sharper softmax than the default initialisation gives
        ln.weight.normal_(1.0, 0.2)
        ln.bias.normal_(0, 0.2)
    x = torch.randn(B * S, D, device='cuda') * 1.5
    want = reference_block(x, B, S, mha, ln)
    W, b, bo = ops.attn_fold_weights(mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, H)
    y = ops.layernorm_split(x, ln.weight.detach(), ln.bias.detach(), eps=ln.eps)
    pl = ops.gemm(y, ops.split_bf16(W), b, want='split')
    got = ops.seq_attention_fold(pl, B, S, H, D, x, bo)
    torch.cuda.synchronize()
    err = float((got.double().cpu() - want).abs().max())
    scale = float(want.abs().max())
    assert err < 3e-5 * max(scale, 1.0), (err, scale)
    # the one-kernel form (LayerNorm + projection + attention + residual, dyg_attn_block)
    got2 = ops.attn_block(x, ln.weight.detach(), ln.bias.detach(), ln.eps, ops.split_bf16(W), b, bo, B, S, H, D)
    torch.cuda.synchronize()
    err2 = float((got2.double().cpu() - want).abs().max())
    assert err2 < 3e-5 * max(scale, 1.0), (err2, scale)


def test_seq_attention_fold_rejects_unsupported():
    x = torch.zeros(130, 200, device='cuda')
    pl = ops.empty_split(130, ops.attn_fold_layout(200, 2)[4], 'cuda')
    with pytest.raises(ValueError):
        ops.seq_attention_fold(pl, 2, 65, 2, 200, x, torch.zeros(200, device='cuda'))
