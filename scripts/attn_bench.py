"""Time the attention sub-block kernels alone at the bench's launch size: python scripts/attn_bench.py [B] [S]"""
import sys
import torch
sys.path.insert(0, '.')
from dyglib_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 12800
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
D, H = 200, 2
torch.manual_seed(0)
mha = torch.nn.MultiheadAttention(D, H).cuda()
ln = torch.nn.LayerNorm(D).cuda()
x = torch.randn(B * S, D, device='cuda')
W, b, bo = ops.attn_fold_weights(mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, H)
Ws = ops.split_bf16(W)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')


def timed(fn, n=5):
    ts = []
    for _ in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts[2:])[len(ts[2:]) // 2], r


t_ln, y = timed(lambda: ops.layernorm_split(x, ln.weight.detach(), ln.bias.detach(), eps=ln.eps))
t_g, pl = timed(lambda: ops.gemm(y, Ws, b, want='split'))
t_a, out = timed(lambda: ops.seq_attention_fold(pl, B, S, H, D, x, bo))
print(f'B={B} S={S}: layernorm_split {t_ln:.3f} ms, projection gemm (N={W.shape[0]}) {t_g:.3f} ms, seq_attention_fold {t_a:.3f} ms')
