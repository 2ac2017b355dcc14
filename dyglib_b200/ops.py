"""Thin torch-tensor wrappers over the C ABI (include/dygb200.h).

torch is used for device memory and streams only; every computation below is one call into
libdygb200.so on the current CUDA stream.  Nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import _native
from ._native import Seg

ACT_NONE, ACT_RELU, ACT_GELU, ACT_SIGMOID = 0, 1, 2, 3

# number of kernels this process launched through the library (bench.py reports it as gpu_launches)
launch_count = 0


def _lib():
    return _native.load()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('dyglib_b200 kernels need CUDA tensors: there is no CPU fallback')
    return ctypes.c_void_p(t.data_ptr())


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise RuntimeError(f'{name} must be a CUDA tensor: dyglib_b200 has no CPU path')
    if t.dtype != dtype:
        raise TypeError(f'{name} must be {dtype}, got {t.dtype}')
    if not t.is_contiguous():
        raise ValueError(f'{name} must be contiguous')
    return t


def _count(n=1):
    global launch_count
    launch_count += n


# Caches of weight-derived operands (BF16x3 planes, folded attention weights, LUTs) are keyed on (data_ptr, tensor version).
# A CUDA-graph replay of a training step updates the parameters without touching their version counters, so every such key also
# carries this epoch; utils.graph.GraphedStep bumps it after each replay of a step captured with grad=True.
WEIGHTS_EPOCH = 0


def bump_weights_epoch():
    global WEIGHTS_EPOCH
    WEIGHTS_EPOCH += 1


# optional per-launch timing (bench.py roofline pass): PROFILE = [] enables it; each entry is
# (kernel name, start event, end event, algorithmic flops, algorithmic bytes)
PROFILE = None


class _Timed:
    def __init__(self, name, flops=0.0, nbytes=0.0):
        self.name, self.flops, self.nbytes = name, flops, nbytes

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if PROFILE is not None:
            self.e1.record()
            PROFILE.append((self.name, self.e0, self.e1, self.flops, self.nbytes))
        return False


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError('dyglib_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')


# ------------------------------------------------------------------ segments for dyg_linear
def seg_rows(table, width=None, idx=None, group=1, table2=None, idx2=None):
    """A gathered / dense row segment: width columns from table[idx] (+ table2[idx2])."""
    s = Seg()
    s.kind = 0
    s.width = int(width if width is not None else table.shape[-1])
    s.group = int(group)
    s.ld = int(table.stride(-2)) if table.dim() >= 2 else s.width
    s.ptr = _p(table).value
    s.idx = _p(idx).value if idx is not None else None
    if table2 is not None:
        s.ptr2 = _p(table2).value
        s.ld2 = int(table2.stride(-2))
        s.idx2 = _p(idx2).value if idx2 is not None else None
    s._keep = (table, idx, table2, idx2)
    return s


def seg_time(dt, w, b, mask_ids=None, group=1, t_query=None, tq_div=1):
    """A time-encoding segment: cos(fma(dt, w, b)), zero where mask_ids == 0.  With ``t_query`` (float64, one
    per ``tq_div`` sub-rows) ``dt`` holds float32 neighbour times and the delta is formed in the kernel."""
    s = Seg()
    s.kind = 1
    s.width = int(w.numel())
    s.group = int(group)
    s.dt = _p(dt).value
    s.w = _p(w).value
    s.b = _p(b).value
    s.mask_ids = _p(mask_ids).value if mask_ids is not None else None
    if t_query is not None:
        s.t_query = _p(t_query).value
        s.tq_div = int(tq_div)
    s._keep = (dt, w, b, mask_ids, t_query)
    return s


# tensor-core dispatch: rows at or above this use dyg_linear_tc (tcgen05 BF16x3); 0 disables it
TC_MIN_ROWS = int(os.environ.get('DYG_TC_MIN_ROWS', 512))
# dense (un-gathered) contractions with at least this many rows go to dyg_gemm_bf16x3 (TMA + tcgen05 CTA pairs)
GEMM_MIN_ROWS = int(os.environ.get('DYG_GEMM_MIN_ROWS', 256))
_tc_weights = {}


def _split_weight(weight, K, ldw):
    """bf16 hi / mid copies of weight[:, :K], zero padded to (n_pad, k_pad); cached per weight version."""
    key = (weight.data_ptr(), weight._version, WEIGHTS_EPOCH, tuple(weight.shape), int(ldw), int(K))
    ent = _tc_weights.get(key)
    if ent is None:
        N = weight.shape[0]
        nt = _lib().dyg_linear_tc_tile(int(N))
        n_pad = (N + nt - 1) // nt * nt
        k_pad = (K + 63) // 64 * 64
        w = torch.as_strided(weight, (N, K), (int(ldw), 1)).float()
        hi = w.to(torch.bfloat16)
        mid = (w - hi.float()).to(torch.bfloat16)
        wh = torch.zeros((n_pad, k_pad), dtype=torch.bfloat16, device=weight.device)
        wm = torch.zeros((n_pad, k_pad), dtype=torch.bfloat16, device=weight.device)
        wh[:N, :K] = hi
        wm[:N, :K] = mid
        if len(_tc_weights) > 256:
            _tc_weights.clear()
        ent = _tc_weights[key] = (wh, wm, n_pad, k_pad, weight)
    return ent


def _tc_ok(segs):
    for s in segs:
        if s.width % 4:
            return False
        if s.kind == 0:
            if s.ld % 4 or (s.ptr or 0) % 16 or (s.ptr2 and (s.ld2 % 4 or s.ptr2 % 16)):
                return False
        elif (s.w or 0) % 16 or (s.b or 0) % 16:
            return False
    return True


def linear(segs, M, weight, bias=None, residual=None, act=ACT_NONE, out=None, out_cols=None,
           c_group=0, c_group_stride=0, c_offset=0, ldw=None, tc=None):
    """out[row(m), :N] = act(A @ weight[:N, :K].T + bias + residual[row(m)])."""
    N = weight.shape[0]
    K = sum(s.width * s.group for s in segs)
    if ldw is None:
        ldw = weight.stride(0)
    if out is None:
        out = torch.empty((M, N), device=weight.device, dtype=torch.float32)
    arr = (Seg * len(segs))(*segs)
    ldc = out.stride(-2)
    ldr = residual.stride(-2) if residual is not None else 0
    if tc is None and GEMM_MIN_ROWS > 0 and M >= GEMM_MIN_ROWS and c_group == 0 and K % 2 == 0 and \
            all(sg.kind == 0 and not sg.idx and not sg.ptr2 and sg.group == 1 and sg.width % 2 == 0 for sg in segs):
        # large dense contraction: BF16x3 operand planes + TMA / tcgen05 CTA-pair GEMM (dyg_gemm_bf16x3)
        a = empty_split(M, K, weight.device)
        col = 0
        for sg in segs:
            split_bf16(sg._keep[0][:, :sg.width], out=a, col=col)
            col += sg.width
        w = weight if (weight.shape[1] == K and weight.stride(0) == ldw) else torch.as_strided(weight, (N, K), (int(ldw), 1))
        return gemm(a, w, bias, residual=residual, act=act, out=out)
    use_tc = (TC_MIN_ROWS > 0 and M >= TC_MIN_ROWS) if tc is None else tc
    if use_tc and _tc_ok(segs):
        wh, wm, n_pad, k_pad, _ = _split_weight(weight, K, ldw)
        with _Timed('linear_tc_kernel', 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N)):
            _native.check(_lib().dyg_linear_tc(arr, len(segs), _p(wh), _p(wm), int(k_pad), int(n_pad), _p(bias), _p(residual),
                                               int(ldr), _p(out), int(ldc), int(M), int(N), int(act), int(c_group),
                                               int(c_group_stride), int(c_offset), _stream()))
        _count()
        return out
    with _Timed('linear_kernel', 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N)):
        _native.check(_lib().dyg_linear(arr, len(segs), _p(weight), int(ldw), _p(bias), _p(residual), int(ldr),
                                        _p(out), int(ldc), int(M), int(N), int(act), int(c_group), int(c_group_stride),
                                        int(c_offset), _stream()))
    _count()
    return out


# ------------------------------------------------------------------ BF16x3 operand pairs + TMA / tcgen05 GEMM
class Split:
    """A fp32 matrix (rows, cols) carried as bf16 planes hi | mid: ``planes`` is a (2, rows, ld) bf16 tensor,
    ld a multiple of 16 so every row starts on a 32-byte boundary (vector stores, TMA strides)."""
    __slots__ = ('planes', 'rows', 'cols')

    def __init__(self, planes, rows, cols):
        self.planes, self.rows, self.cols = planes, rows, cols

    @property
    def ld(self):
        return self.planes.shape[2]

    @property
    def hi(self):
        return self.planes[0]

    @property
    def mid(self):
        return self.planes[1]

    def float(self):
        return (self.planes[0].float() + self.planes[1].float())[:, :self.cols]


def empty_split(rows, cols, device):
    ld = (cols + 15) // 16 * 16
    return Split(torch.empty((2, rows, ld), dtype=torch.bfloat16, device=device), rows, cols)


def split_bf16(x, out=None, col=0):
    """hi | mid planes of a fp32 matrix (last dimension contiguous); with ``out`` / ``col`` the planes are written into
    the column window [col, col + D) of an existing Split (concatenation without a copy)."""
    M, D = x.shape
    if out is None:
        out = empty_split(M, D, x.device)
    hi, mid = out.hi, out.mid
    if col:
        hi, mid = hi[:, col:], mid[:, col:]
    _native.check(_lib().dyg_split_bf16(_p(x), int(x.stride(0)), int(M), int(D), ctypes.c_void_p(hi.data_ptr()),
                                        ctypes.c_void_p(mid.data_ptr()), int(out.ld), _stream()))
    _count()
    return out


def mlp2(xs, w1, b1, w2, b2, act2=ACT_NONE):
    """act2(W2 relu(W1 cat(xs) + b1) + b2) for dense fp32 row blocks ``xs`` (MergeLayer, models/modules.py:57-68).  With enough
    rows for the tcgen05 GEMM the hidden activation leaves the first GEMM's epilogue as BF16x3 operand planes, so the second
    GEMM needs no split launch (one launch less in steps that are launch-bound)."""
    M = xs[0].shape[0]
    K = sum(x.shape[1] for x in xs)
    if GEMM_MIN_ROWS > 0 and M >= GEMM_MIN_ROWS and all(x.shape[1] % 2 == 0 for x in xs):
        a = empty_split(M, K, xs[0].device)
        col = 0
        for x in xs:
            split_bf16(x, out=a, col=col)
            col += x.shape[1]
        h = gemm(a, w1, b1, act=ACT_RELU, want='split')
        return gemm(h, w2, b2, act=act2)
    h = linear([seg_rows(x) for x in xs], M, w1, b1, act=ACT_RELU)
    return linear([seg_rows(h)], M, w2, b2, act=act2)


_split_weights = {}


def split_weight(weight):
    """Cached BF16x3 planes of a (N, K) fp32 weight (re-split when the parameter changes)."""
    key = (weight.data_ptr(), weight._version, WEIGHTS_EPOCH, tuple(weight.shape), tuple(weight.stride()))
    ent = _split_weights.get(key)
    if ent is None:
        if len(_split_weights) > 256:
            _split_weights.clear()
        w = weight.detach()
        if w.stride(-1) != 1:
            w = w.contiguous()
        ent = _split_weights[key] = (split_bf16(w.float()), weight)
    return ent[0]


def gemm(a, weight, bias=None, residual=None, act=ACT_NONE, out=None, out_split=None, want='f32'):
    """act(a @ weight.T + bias + residual) on tcgen05 (dyg_gemm_bf16x3).  ``a``: Split; ``weight``: fp32 (N, K) parameter
    (split once and cached) or a Split.  ``want``: 'f32' -> fp32 tensor, 'split' -> Split, 'both' -> (fp32, Split)."""
    w = weight if isinstance(weight, Split) else split_weight(weight)
    M, K, N = a.rows, a.cols, w.rows
    if w.cols != K:
        raise ValueError(f'gemm: A has {K} columns, W has {w.cols}')
    dev = a.planes.device
    if want in ('f32', 'both') and out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=dev)
    if want in ('split', 'both') and out_split is None:
        out_split = empty_split(M, N, dev)
    if want == 'f32':
        out_split = None
    if want == 'split':
        out = None
    with _Timed('gemm_bf16x3_kernel', 2.0 * M * N * K, 4.0 * (M * K + N * K + M * N) * (2 if want == 'both' else 1)):
        _native.check(_lib().dyg_gemm_bf16x3(
            _p(a.hi), _p(a.mid), int(a.ld), _p(w.hi), _p(w.mid), int(w.ld), _p(bias), _p(residual),
            int(residual.stride(0)) if residual is not None else 0, _p(out), int(out.stride(0)) if out is not None else 0,
            _p(out_split.hi) if out_split is not None else None, _p(out_split.mid) if out_split is not None else None,
            int(out_split.ld) if out_split is not None else 0, int(M), int(N), int(K), int(act), _stream()))
    _count()
    if want == 'f32':
        return out
    if want == 'split':
        return out_split
    return out, out_split


def table_planes(table):
    """BF16x3 operand planes of a fp32 feature table, rows padded with zeros to a multiple of 16 columns (the layout
    dyg_patch_project gathers from)."""
    t = table.detach().float().contiguous()
    out = empty_split(t.shape[0], t.shape[1], t.device)
    out.planes.zero_()
    return split_bf16(t, out)


def patch_project_stage_blocks(F_node, F_edge, T, F_lut, P):
    """(number of stages, 32-column blocks per (type, p)) of dyg_patch_project for these widths."""
    nblk = (ctypes.c_int32 * 5)()
    nst = _lib().dyg_patch_project_stages(int(F_node), int(F_edge), int(T), int(F_lut), int(P), nblk)
    return nst, list(nblk)


def pack_patch_weights(w_node, w_edge, w_time, w_cooc, P):
    """Packed BF16x3 weight planes (2, 64, 32*stages) for dyg_patch_project: stage s = (type, p, block) owns columns
    [32 s, 32 s + 32) holding W_type[:, p*F + 32*block : p*F + min(F, 32*block + 32)] (zero elsewhere); the two
    co-occurrence LUT gathers (count in src row, count in dst row) both use the co-occurrence weight."""
    C = w_node.shape[0]
    ws = [w_node, w_edge, w_time, w_cooc, w_cooc]
    Fs = [w.shape[1] // P for w in ws]
    nst, nblk = patch_project_stage_blocks(Fs[0], Fs[1], Fs[2], Fs[3], P)
    dev = w_node.device
    packed = torch.zeros((64, nst * 32), dtype=torch.float32, device=dev)
    s = 0
    for ty in range(5):
        w, F = ws[ty].detach().float(), Fs[ty]
        for p in range(P):
            for blk in range(nblk[ty]):
                lo, hi = blk * 32, min(F, blk * 32 + 32)
                if hi > lo:
                    packed[:C, s * 32:s * 32 + hi - lo] = w[:, p * F + lo:p * F + hi]
                s += 1
    assert s == nst
    return split_bf16(packed)


def patch_project(sides, node_planes, F_node, edge_planes, F_edge, lut_planes, F_lut, t_query, tw, tb, packed_w, bias, P, C, S, X,
                  zero_rows=0):
    """dyg_patch_project.  ``sides``: list of (ids, eids, t_nbr, cnt_a, cnt_b, ntok, tok_off) with (B, Lp) device tensors.
    ``zero_rows``: bit 0 / 1 = row 0 of the node / edge table is all zero (padded positions skip the gather)."""
    arr = (_native.ProjSide * len(sides))()
    keep = []
    rows = 0
    for i, (ids, eids, tn, ca, cb, ntok, off) in enumerate(sides):
        B = ids.shape[0]
        for t_, dt_ in ((ids, torch.int64), (eids, torch.int64), (tn, torch.float32), (ca, torch.int64), (cb, torch.int64)):
            _chk(t_, dt_, 'patch_project side array')
        arr[i].ids, arr[i].eids, arr[i].t_nbr = _p(ids).value, _p(eids).value, _p(tn).value
        arr[i].cnt_a, arr[i].cnt_b = _p(ca).value, _p(cb).value
        arr[i].tokens, arr[i].ntok, arr[i].tok_off = int(B * ntok), int(ntok), int(off)
        rows += B * ntok
        keep.append((ids, eids, tn, ca, cb))
    T = tw.numel()
    K = P * (F_node + F_edge + T + 2 * F_lut)
    with _Timed('patch_project_kernel', 2.0 * rows * C * K, rows * (P * (2.0 * 4 * (F_node + F_edge) + 8.0 * F_lut + 28.0) + 16.0 * C)):
        _native.check(_lib().dyg_patch_project(
            arr, len(sides), _p(node_planes.hi), _p(node_planes.mid), int(node_planes.ld), int(F_node),
            _p(edge_planes.hi), _p(edge_planes.mid), int(edge_planes.ld), int(F_edge),
            _p(lut_planes.hi), _p(lut_planes.mid), int(lut_planes.ld), int(F_lut), int(zero_rows), _p(t_query), _p(tw), _p(tb), int(T),
            _p(packed_w.hi), _p(packed_w.mid), int(packed_w.ld), _p(bias), int(P), int(C), int(S), _p(X), int(X.stride(0)),
            _stream()))
    _count()
    return X


_ffn_ws = {}


def _ffn_workspace(device):
    """Per-device, per-stream scratch of dyg_ln_ffn_bf16x3 / dyg_ln_gemm_bf16x3 (one A-operand image per CTA, stays in L2)."""
    # one scratch per (device, stream): two forwards on different streams of a device must not share the operand images
    key = (str(device), int(torch.cuda.current_stream(device).cuda_stream))
    if key not in _ffn_ws:
        _ffn_ws[key] = torch.empty(int(_lib().dyg_ln_ffn_workspace_bytes()), dtype=torch.uint8, device=device)
    return _ffn_ws[key]


def ffn_fusable(D, Dff):
    return D % 8 == 0 and D <= 208 and Dff % 32 == 0


def ln_ffn(x, gamma, beta, eps, w1, b1, w2, b2, out=None):
    """x + W2 gelu(W1 LayerNorm(x) + b1) + b2 in one kernel (dyg_ln_ffn_bf16x3); the hidden activation never reaches HBM."""
    M, D = x.shape
    Dff = w1.shape[0]
    if out is None:
        out = torch.empty_like(x)
    w1s, w2s = split_weight(w1), split_weight(w2)
    ws = _ffn_workspace(x.device)
    with _Timed('ln_ffn_bf16x3_kernel', 4.0 * M * D * Dff, 8.0 * M * D):
        _native.check(_lib().dyg_ln_ffn_bf16x3(_p(x), int(x.stride(0)), _p(gamma), _p(beta), float(eps), _p(w1s.hi), _p(w1s.mid),
                                               int(w1s.ld), _p(b1), _p(w2s.hi), _p(w2s.mid), int(w2s.ld), _p(b2), _p(out), int(out.stride(0)),
                                               int(M), int(D), int(Dff), _p(ws), int(ws.numel()), _stream()))
    _count()
    return out


def ln_gemm(x, gamma, beta, eps, weight, bias=None, act=ACT_NONE, want='split'):
    """act(LayerNorm(x) @ weight.T + bias) in one kernel (dyg_ln_gemm_bf16x3): the normalised rows never reach HBM.
    ``weight``: fp32 (N, D) parameter or Split; ``want`` as in ``gemm``."""
    w = weight if isinstance(weight, Split) else split_weight(weight)
    M, D = x.shape
    N = w.rows
    if w.cols != D:
        raise ValueError(f'ln_gemm: x has {D} columns, W has {w.cols}')
    dev = x.device
    out = torch.empty((M, N), dtype=torch.float32, device=dev) if want in ('f32', 'both') else None
    sp = empty_split(M, N, dev) if want in ('split', 'both') else None
    ws = _ffn_workspace(dev)
    with _Timed('ln_gemm_bf16x3_kernel', 2.0 * M * N * D, 4.0 * (M * D + N * D + M * N) * (2 if want == 'both' else 1)):
        _native.check(_lib().dyg_ln_gemm_bf16x3(
            _p(x), int(x.stride(0)), _p(gamma), _p(beta), float(eps), _p(w.hi), _p(w.mid), int(w.ld), _p(bias),
            _p(out), int(out.stride(0)) if out is not None else 0, _p(sp.hi) if sp is not None else None,
            _p(sp.mid) if sp is not None else None, int(sp.ld) if sp is not None else 0, int(M), int(N), int(D), int(act),
            _p(ws), int(ws.numel()), _stream()))
    _count()
    return out if want == 'f32' else sp if want == 'split' else (out, sp)


def ln_gemm_fusable(D):
    return D % 8 == 0 and D <= 224


def layernorm_split(x, gamma, beta, eps=1e-5, out=None, y=None):
    """LayerNorm(x) as a Split (and optionally also as fp32 ``y``)."""
    M, D = x.shape
    if out is None:
        out = empty_split(M, D, x.device)
    with _Timed('layernorm_split_kernel', 8.0 * M * D, 8.0 * M * D):
        _native.check(_lib().dyg_layernorm_split(_p(x), int(x.stride(0)), _p(gamma), _p(beta), float(eps), _p(y),
                                                 int(y.stride(0)) if y is not None else 0, _p(out.hi), _p(out.mid), int(out.ld),
                                                 int(M), int(D), _stream()))
    _count()
    return out


def layernorm(x, gamma, beta, r1=None, F1=0, rconst=None, eps=1e-5, out=None):
    M, D = x.shape
    if out is None:
        out = torch.empty_like(x)
    _native.check(_lib().dyg_layernorm(_p(x), x.stride(0), _p(r1), r1.stride(0) if r1 is not None else 0,
                                       int(F1 if (r1 is not None or rconst is not None) else D), _p(rconst),
                                       _p(gamma), _p(beta), float(eps), _p(out), out.stride(0), int(M), int(D), _stream()))
    _count()
    return out


def gather_rows(table, idx, table2=None, out=None):
    M = idx.numel()
    D = table.shape[1]
    if out is None:
        out = torch.empty((M, D), device=table.device, dtype=torch.float32)
    _native.check(_lib().dyg_gather_rows(_p(table), table.stride(0), _p(table2),
                                         table2.stride(0) if table2 is not None else 0, _p(idx), int(M), int(D),
                                         _p(out), out.stride(0), _stream()))
    _count()
    return out


def time_encode(dt, w, b, out=None):
    n = dt.numel()
    T = w.numel()
    if out is None:
        out = torch.empty((n, T), device=dt.device, dtype=torch.float32)
    _native.check(_lib().dyg_time_encode(_p(dt), int(n), _p(w), _p(b), int(T), _p(out), _stream()))
    _count()
    return out


def time_encode_bwd(dt, w, b, grad_out):
    """(grad_w (T,), grad_b (T,)) of dyg_time_encode for ``grad_out`` (n, T)."""
    n, T = dt.numel(), w.numel()
    gw = torch.zeros(T, device=dt.device, dtype=torch.float32)
    gb = torch.zeros(T, device=dt.device, dtype=torch.float32)
    _native.check(_lib().dyg_time_encode_bwd(_p(dt), int(n), _p(w), _p(b), int(T), _p(grad_out), int(grad_out.stride(0)), _p(gw), _p(gb),
                                             _stream()))
    _count()
    return gw, gb


def zero_row0_flags(node_tab, edge_tab, node_tab2=None):
    """Which padding rows (row 0: node 0 / edge 0, preprocess_data/preprocess_data.py:101-108) of static tables are all
    zeros, so that dyg_temporal_attend may skip reading them.  One host sync: call it at model construction."""
    def z(t):
        return t is None or bool(torch.count_nonzero(t[0]).item() == 0)
    return (1 if (z(node_tab) and z(node_tab2)) else 0) | (2 if z(edge_tab) else 0)


def temporal_attend(qk, n, k, H, node_tab, node_idx, F, edge_tab, edge_idx, E, T, mask_ids, node_tab2=None,
                    time_feat=None, t_query=None, t_nbr=None, w=None, b=None, want_scores=False, zero_row0=0,
                    prob_scale=None):
    """zero_row0: see zero_row0_flags (bit 0: row 0 of node_tab / node_tab2 is zero, bit 1: row 0 of edge_tab is zero)."""
    Dk = F + E + T
    out = torch.empty((n, H * Dk), device=qk.device, dtype=torch.float32)
    scores = torch.empty((n, H, k), device=qk.device, dtype=torch.float32) if want_scores else None
    rows = 4.0 * (F * (2 if node_tab2 is not None else 1) + E + (T if time_feat is not None else 0))
    with _Timed('temporal_attend_kernel', 4.0 * n * k * H * Dk, n * (k * (rows + 28.0) + 8.0 * H * Dk)):
        _native.check(_lib().dyg_temporal_attend(
            _p(qk), qk.stride(0), int(n), int(k), int(H), _p(node_tab), node_tab.stride(-2), _p(node_tab2),
            node_tab2.stride(-2) if node_tab2 is not None else 0, _p(node_idx), int(F), _p(edge_tab), edge_tab.stride(-2),
            _p(edge_idx), int(E), _p(time_feat), _p(t_query), _p(t_nbr), _p(w), _p(b), int(T), _p(mask_ids),
            _p(out), out.stride(0), _p(scores), int(zero_row0), _p(prob_scale), _stream()))
    _count()
    return out, scores


def temporal_attend_bwd(qk, n, k, H, node_tab, node_idx, F, edge_tab, edge_idx, E, T, mask_ids, t_query, t_nbr, w, b, probs,
                        prob_scale, s_out, grad_s, want_nbr=False, node_tab2=None):
    """Backward of ``temporal_attend`` (dyg_temporal_attend_bwd): returns (grad_qk, grad_nbr or None, grad_w, grad_b)."""
    Dk = F + E + T
    dev = qk.device
    gqk = torch.empty((n, H * Dk), device=dev, dtype=torch.float32)
    gnbr = torch.empty((n * k, F), device=dev, dtype=torch.float32) if want_nbr else None
    gw = torch.zeros(T, device=dev, dtype=torch.float32)
    gb = torch.zeros(T, device=dev, dtype=torch.float32)
    _native.check(_lib().dyg_temporal_attend_bwd(
        _p(qk), qk.stride(0), int(n), int(k), int(H), _p(node_tab), node_tab.stride(-2), _p(node_tab2),
        node_tab2.stride(-2) if node_tab2 is not None else 0, _p(node_idx), int(F), _p(edge_tab), edge_tab.stride(-2),
        _p(edge_idx), int(E), _p(t_query), _p(t_nbr), _p(w), _p(b), int(T), _p(mask_ids), _p(probs), _p(prob_scale),
        _p(s_out), s_out.stride(0), _p(grad_s), grad_s.stride(0), _p(gqk), gqk.stride(0), _p(gnbr),
        gnbr.stride(0) if gnbr is not None else 0, _p(gw), _p(gb), _stream()))
    _count()
    return gqk, gnbr, gw, gb


def seq_attention(qkv, B, S, H, hd, out=None):
    if out is None:
        out = torch.empty((B * S, H * hd), device=qkv.device, dtype=torch.float32)
    with _Timed('seq_attention_kernel', 4.0 * B * H * S * S * hd, 16.0 * B * S * H * hd):
        _native.check(_lib().dyg_seq_attention(_p(qkv), qkv.stride(-2), int(B), int(S), int(H), int(hd), _p(out),
                                               out.stride(-2), _stream()))
    _count()
    return out


def seq_attention_tc(qkv, B, S, H, hd, want='split'):
    """Tensor-core sequence attention (S <= 128, even hd <= 128).  want: 'f32' | 'split' | 'both'."""
    dev = qkv.device
    out = torch.empty((B * S, H * hd), device=dev, dtype=torch.float32) if want in ('f32', 'both') else None
    sp = empty_split(B * S, H * hd, dev) if want in ('split', 'both') else None
    with _Timed('seq_attention_mma_kernel', 4.0 * B * H * S * S * hd, 16.0 * B * S * H * hd):
        _native.check(_lib().dyg_seq_attention_tc(_p(qkv), qkv.stride(-2), int(B), int(S), int(H), int(hd), _p(out),
                                                  out.stride(-2) if out is not None else 0,
                                                  _p(sp.hi) if sp is not None else None, _p(sp.mid) if sp is not None else None,
                                                  int(sp.ld) if sp is not None else 0, _stream()))
    _count()
    return out if want == 'f32' else sp if want == 'split' else (out, sp)


def attn_fold_layout(D, H):
    """Column layout of the projection planes dyg_seq_attention_fold reads: (q_col0, k_col0, v_col0, hdk, total columns)."""
    hd = D // H
    hdk = (hd + 7) // 8 * 8
    k0 = H * hdk
    v0 = k0 + H * hdk
    return 0, k0, v0, hdk, v0 + H * D


def attn_fold_fusable(S, D, H):
    hd = D // H
    return S <= 64 and D % H == 0 and hd % 4 == 0 and hd <= 112 and D % 8 == 0 and D <= 208


def attn_fold_weights(in_proj_weight, in_proj_bias, out_proj_weight, out_proj_bias, H):
    """(W_cat (N, D) fp32, b_cat (N,), b_out (D,)) of the folded attention block: rows [q * log2(e)/sqrt(hd) | k | W_o,h W_v,h],
    computed in float64.  softmax rows sum to one, so W_o,h b_v,h rides inside v' and b_o is added once at the end."""
    D = in_proj_weight.shape[1]
    hd = D // H
    q0, k0, v0, hdk, N = attn_fold_layout(D, H)
    w = in_proj_weight.detach().double()
    b = in_proj_bias.detach().double()
    wo = out_proj_weight.detach().double()
    W = torch.zeros((N, D), dtype=torch.float64, device=w.device)
    bb = torch.zeros(N, dtype=torch.float64, device=w.device)
    c = 1.4426950408889634 / (hd ** 0.5)
    for h in range(H):
        sl = slice(h * hd, (h + 1) * hd)
        W[q0 + h * hdk:q0 + h * hdk + hd] = w[:D][sl] * c
        bb[q0 + h * hdk:q0 + h * hdk + hd] = b[:D][sl] * c
        W[k0 + h * hdk:k0 + h * hdk + hd] = w[D:2 * D][sl]
        bb[k0 + h * hdk:k0 + h * hdk + hd] = b[D:2 * D][sl]
        W[v0 + h * D:v0 + (h + 1) * D] = wo[:, sl] @ w[2 * D:][sl]
        bb[v0 + h * D:v0 + (h + 1) * D] = wo[:, sl] @ b[2 * D:][sl]
    return W.float().contiguous(), bb.float().contiguous(), out_proj_bias.detach().float().contiguous()


def seq_attention_fold(planes, B, S, H, D, x, bias, out=None):
    """x + bias + sum_h softmax(q_h k_h^T) v'_h on tcgen05 (dyg_seq_attention_fold); ``planes``: Split (B*S, attn_fold_layout(D, H)[4])
    written by the projection GEMM with attn_fold_weights."""
    q0, k0, v0, hdk, N = attn_fold_layout(D, H)
    hd = D // H
    if planes.cols != N or planes.rows != B * S:
        raise ValueError(f'seq_attention_fold: planes are {planes.rows} x {planes.cols}, expected {B * S} x {N}')
    if out is None:
        out = torch.empty((B * S, D), device=x.device, dtype=torch.float32)
    with _Timed('seq_attention_fold_kernel', 2.0 * B * H * S * S * (hd + D), 4.0 * B * S * (N + 2 * D)):
        _native.check(_lib().dyg_seq_attention_fold(_p(planes.hi), _p(planes.mid), int(planes.ld), q0, k0, v0, int(B), int(S), int(H),
                                                    int(hd), int(D), _p(x), int(x.stride(0)), _p(bias), _p(out), int(out.stride(0)),
                                                    _stream()))
    _count()
    return out


_attn_ws = {}


def attn_block(x, gamma, beta, eps, wcat, bcat, bout, B, S, H, D, out=None):
    """x + bout + sum_h softmax(q_h k_h^T) v'_h with [q | k | v'] = wcat LayerNorm(x) + bcat, one kernel (dyg_attn_block):
    LayerNorm, projection, attention and residual; the projected rows live in an L2-resident scratch.  ``wcat``: Split of
    attn_fold_weights(...)[0]."""
    q0, k0, v0, hdk, N = attn_fold_layout(D, H)
    if wcat.rows != N or wcat.cols != D or x.shape != (B * S, D):
        raise ValueError(f'attn_block: weight {wcat.rows} x {wcat.cols} / x {tuple(x.shape)} do not match B={B} S={S} D={D} H={H}')
    if out is None:
        out = torch.empty_like(x)
    key = (str(x.device), int(torch.cuda.current_stream(x.device).cuda_stream), N)
    if key not in _attn_ws:
        _attn_ws[key] = torch.empty(int(_lib().dyg_attn_block_workspace_bytes(int(N))) + 1024, dtype=torch.uint8, device=x.device)
    ws = _attn_ws[key]
    off = (-ws.data_ptr()) % 1024
    hd = D // H
    with _Timed('attn_block_kernel', 2.0 * B * S * N * D + 2.0 * B * H * S * S * (hd + D), 8.0 * B * S * D):
        _native.check(_lib().dyg_attn_block(_p(x), int(x.stride(0)), _p(gamma), _p(beta), float(eps), _p(wcat.hi), _p(wcat.mid), int(wcat.ld),
                                            _p(bcat), int(N), q0, k0, v0, _p(bout), int(B), int(S), int(H), int(hd), int(D), _p(out),
                                            int(out.stride(0)), ctypes.c_void_p(ws.data_ptr() + off), int(ws.numel() - off), _stream()))
    _count()
    return out


# ------------------------------------------------------------------ stable radix sort (csrc/sort.cu): the CSR build's ordering
def stable_argsort(keys):
    """Positions that sort ``keys`` stably (LSD radix sort on the raw bit pattern, 8 bits per pass; passes whose digit is the same for
    every key are skipped).  ``keys``: int32 / int64 CUDA tensor holding UNSIGNED-orderable bit patterns.  Returns int64 positions."""
    n = keys.numel()
    kb = keys.element_size()
    if kb not in (4, 8):
        raise TypeError('stable_argsort: 4- or 8-byte keys')
    dev = keys.device
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=dev)
    keys = keys.contiguous()
    hist = torch.empty(kb * 256, dtype=torch.int64, device=dev)
    _native.check(_lib().dyg_radix_digit_hist(_p(keys), kb, int(n), _p(hist), _stream()))
    _count()
    h = hist.cpu().reshape(kb, 256)
    passes = [p for p in range(kb) if int(h[p].max()) < n]
    if not passes:
        return torch.arange(n, dtype=torch.int64, device=dev)
    ws = torch.empty(int(_lib().dyg_radix_sort_workspace_entries(int(n))), dtype=torch.int32, device=dev)
    ka, kbuf = keys, torch.empty_like(keys)
    va, vb = None, torch.empty(n, dtype=torch.int32, device=dev)
    spare = torch.empty(n, dtype=torch.int32, device=dev) if len(passes) > 1 else None
    for p in passes:
        _native.check(_lib().dyg_radix_sort_pass(_p(ka), _p(va), _p(kbuf), _p(vb), kb, int(n), int(p), _p(hist), _p(ws), _stream()))
        _count(3)
        if ka is keys:
            ka, kbuf = kbuf, torch.empty_like(keys)
        else:
            ka, kbuf = kbuf, ka
        va, vb = vb, (spare if va is None else va)
    return va.to(torch.int64) & 0xFFFFFFFF


def float64_sort_key(t):
    """Bit pattern of float64 values whose UNSIGNED order equals the numeric order (sign bit flipped for non-negatives, all bits for
    negatives); -0.0 sorts before +0.0, which a stable sort by value would keep in input order: callers must not mix them."""
    b = t.contiguous().view(torch.int64)
    return b ^ ((b >> 63) | torch.iinfo(torch.int64).min)


# ------------------------------------------------------------------ training path (backward kernels, csrc/train.cu)
# input gradients of dense layers above this many flops go to the tcgen05 GEMM (three more launches: split, transpose, split);
# below it one BF16x3 mma.sync launch (dyg_gemm_dx)
GEMM_DX_TC_FLOPS = float(os.environ.get('DYG_GEMM_DX_TC_FLOPS', '3e9'))
# dense layers below this many flops (2 M N K) take the one-launch backward (dyg_linear_bwd)
LINEAR_BWD_FUSED_FLOPS = float(os.environ.get('DYG_LINEAR_BWD_FUSED_FLOPS', '2e8'))
# weight gradients above this many flops: planes of g^T and x^T through the tcgen05 GEMM (four more launches)
GEMM_DW_TC_FLOPS = float(os.environ.get('DYG_GEMM_DW_TC_FLOPS', '1.2e10'))


def gemm_dw(g, x, dw=None, db=None, want_bias=False):
    """(dW (N, K), db (N) or None) of y = x W^T + b from g = dL/dy (M, N): dW = g^T x ACCUMULATED into ``dw`` (zeroed when created
    here).  Small layers: one split-over-rows fp32 launch (dyg_gemm_dw).  Large layers: the tcgen05 GEMM on planes of g^T and
    x^T (contraction over the rows), bias gradient from dyg_gemm_dw against a column of ones."""
    M, N = g.shape
    K = x.shape[1]
    fresh = dw is None
    if fresh:
        dw = torch.zeros((N, K), dtype=torch.float32, device=g.device)
    if want_bias and db is None:
        db = torch.zeros(N, dtype=torch.float32, device=g.device)
    if 2.0 * M * N * K >= GEMM_DW_TC_FLOPS:
        part = gemm(split_bf16(g.t().contiguous()), split_bf16(x.t().contiguous().float()))
        if fresh:
            dw = part
        else:
            dw += part
        if db is not None:
            ones = torch.ones((M, 1), dtype=torch.float32, device=g.device)
            scratch = torch.zeros((N, 1), dtype=torch.float32, device=g.device)
            _native.check(_lib().dyg_gemm_dw(_p(g), int(g.stride(0)), _p(ones), 1, int(M), int(N), 1, _p(scratch), 1, _p(db), _stream()))
            _count()
        return dw, db
    _native.check(_lib().dyg_gemm_dw(_p(g), int(g.stride(0)), _p(x), int(x.stride(0)), int(M), int(N), int(K), _p(dw), int(dw.stride(0)),
                                     _p(db), _stream()))
    _count()
    return dw, db


def linear_bwd(g, y_mask, xs, widths, w, need_x=True, need_w=True, need_b=False):
    """Backward of a small dense layer y = act(cat(xs) W^T + b) in one launch per input segment (dyg_linear_bwd):
    (dX (M, sum widths) or None, dW (N, sum widths) or None, db (N) or None); ``y_mask``: the ReLU output (masks g) or None."""
    M, N = g.shape
    Kt = sum(widths)
    dev = g.device
    dx = torch.empty((M, Kt), dtype=torch.float32, device=dev) if need_x else None
    dw = db = None
    if need_w or need_b:
        buf = torch.zeros(N * (Kt + 1), dtype=torch.float32, device=dev)                       # dW | db: one memset
        dw = buf[:N * Kt].view(N, Kt)
        db = buf[N * Kt:] if need_b else None
    off = 0
    for i, x in enumerate(xs):
        K = widths[i]
        _native.check(_lib().dyg_linear_bwd(
            _p(g), int(g.stride(0)), _p(y_mask), int(y_mask.stride(0)) if y_mask is not None else 0,
            _p(x) if dw is not None else None, int(x.stride(0)) if dw is not None else 0,
            ctypes.c_void_p(w.data_ptr() + 4 * off) if need_x else None, int(w.stride(0)), int(M), int(N), int(K),
            ctypes.c_void_p(dx.data_ptr() + 4 * off) if need_x else None, Kt,
            ctypes.c_void_p(dw.data_ptr() + 4 * off) if dw is not None else None, Kt, _p(db) if (db is not None and i == 0) else None, _stream()))
        _count()
        off += K
    return dx, (dw if need_w else None), db


def gemm_dx(g, w, y_mask=None):
    """dX = (g * (y_mask > 0)) @ w for y = act(x W^T): one BF16x3 mma.sync launch (dyg_gemm_dx); above GEMM_DX_TC_FLOPS the tcgen05
    GEMM on planes of g and W^T."""
    M, N = g.shape
    K = w.shape[1]
    if 2.0 * M * N * K >= GEMM_DX_TC_FLOPS:
        if y_mask is not None:
            g = g * (y_mask > 0)
        return gemm(split_bf16(g), split_bf16(w.t().contiguous().float()))
    dx = torch.empty((M, K), dtype=torch.float32, device=g.device)
    _native.check(_lib().dyg_gemm_dx(_p(g), int(g.stride(0)), _p(y_mask), int(y_mask.stride(0)) if y_mask is not None else 0, _p(w),
                                     int(w.stride(0)), int(M), int(N), int(K), _p(dx), int(dx.stride(0)), _stream()))
    _count()
    return dx


def layernorm_bwd(x, gamma, eps, dy):
    """(dx, dgamma, dbeta) of LayerNorm(x) * gamma + beta (dyg_layernorm_bwd)."""
    M, D = x.shape
    dx = torch.empty((M, D), dtype=torch.float32, device=x.device)
    dg = torch.zeros(D, dtype=torch.float32, device=x.device)
    dbt = torch.zeros(D, dtype=torch.float32, device=x.device)
    _native.check(_lib().dyg_layernorm_bwd(_p(x), int(x.stride(0)), _p(gamma), float(eps), _p(dy), int(dy.stride(0)), _p(dx), int(dx.stride(0)),
                                           _p(dg), _p(dbt), int(M), int(D), _stream()))
    _count()
    return dx, dg, dbt


def gelu_fwd(v, mask=None, want='f32'):
    """gelu(v) * mask as fp32 ('f32'), operand planes ('split') or both."""
    M, N = v.shape
    h = torch.empty((M, N), dtype=torch.float32, device=v.device) if want in ('f32', 'both') else None
    sp = empty_split(M, N, v.device) if want in ('split', 'both') else None
    _native.check(_lib().dyg_gelu_fwd(_p(v), int(v.stride(0)), _p(mask), int(mask.stride(0)) if mask is not None else 0, _p(h),
                                      int(h.stride(0)) if h is not None else 0, _p(sp.hi) if sp is not None else None,
                                      _p(sp.mid) if sp is not None else None, int(sp.ld) if sp is not None else 0, int(M), int(N), _stream()))
    _count()
    return h if want == 'f32' else sp if want == 'split' else (h, sp)


def gelu_bwd(v, mask, dh):
    M, N = v.shape
    dv = torch.empty((M, N), dtype=torch.float32, device=v.device)
    _native.check(_lib().dyg_gelu_bwd(_p(v), int(v.stride(0)), _p(mask), int(mask.stride(0)) if mask is not None else 0, _p(dh),
                                      int(dh.stride(0)), _p(dv), int(dv.stride(0)), int(M), int(N), _stream()))
    _count()
    return dv


def seq_attention_train_fwd(qkv, B, S, H, hd, prob_mask=None):
    """(out (B*S, H*hd), probs (B, H, S, S)) with the softmax kept for the backward pass (dyg_seq_attention_train_fwd)."""
    out = torch.empty((B * S, H * hd), dtype=torch.float32, device=qkv.device)
    probs = torch.empty((B, H, S, S), dtype=torch.float32, device=qkv.device)
    _native.check(_lib().dyg_seq_attention_train_fwd(_p(qkv), int(qkv.stride(0)), int(B), int(S), int(H), int(hd), _p(prob_mask), _p(probs),
                                                     _p(out), int(out.stride(0)), _stream()))
    _count()
    return out, probs


def seq_attention_train_bwd(qkv, B, S, H, hd, prob_mask, probs, dout):
    dqkv = torch.empty((B * S, 3 * H * hd), dtype=torch.float32, device=qkv.device)
    _native.check(_lib().dyg_seq_attention_train_bwd(_p(qkv), int(qkv.stride(0)), int(B), int(S), int(H), int(hd), _p(prob_mask), _p(probs),
                                                     _p(dout), int(dout.stride(0)), _p(dqkv), int(dqkv.stride(0)), _stream()))
    _count()
    return dqkv


def mean_tokens(x, B, S, D, tok0, cnt, out=None):
    if out is None:
        out = torch.empty((B, D), device=x.device, dtype=torch.float32)
    _native.check(_lib().dyg_mean_tokens(_p(x), int(B), int(S), int(D), int(tok0), int(cnt), _p(out), out.stride(0), _stream()))
    _count()
    return out


def cooc_count(src_ids, dst_ids, want_float=True, want_int=False):
    """src_ids (B,Ls), dst_ids (B,Ld) int64 device tensors."""
    B, Ls = src_ids.shape
    Ld = dst_ids.shape[1]
    dev = src_ids.device
    fs = torch.empty((B, Ls, 2), device=dev, dtype=torch.float32) if want_float else None
    fd = torch.empty((B, Ld, 2), device=dev, dtype=torch.float32) if want_float else None
    cs = torch.empty((2, B, Ls), device=dev, dtype=torch.int64) if want_int else None
    cd = torch.empty((2, B, Ld), device=dev, dtype=torch.int64) if want_int else None
    _native.check(_lib().dyg_cooc_count(_p(src_ids), src_ids.stride(0), _p(dst_ids), dst_ids.stride(0), int(B), int(Ls),
                                        int(Ld), _p(fs), _p(fd), _p(cs), _p(cd), _stream()))
    _count()
    return fs, fd, cs, cd
