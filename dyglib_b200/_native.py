"""Loader for libdygb200.so, the C-ABI library declared in include/dygb200.h.

There is no CPU fallback: if the shared library is missing and cannot be built, or no CUDA
device is present when a kernel is requested, the call raises.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, 'csrc')
_INCLUDE = os.path.join(os.path.dirname(_HERE), 'include')
LIB_PATH = os.path.join(_HERE, 'libdygb200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC']

# DYG_ABI_VERSION of include/dygb200.h these SIGNATURES were written against (bumped with every prototype change)
ABI_VERSION = 13

_lock = threading.Lock()
_lib = None


def sources():
    return sorted(glob.glob(os.path.join(_CSRC, '*.cu')))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(_CSRC, '*.cuh')) + glob.glob(os.path.join(_INCLUDE, '*.h'))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into dyglib_b200/libdygb200.so (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get('NVCC', 'nvcc')
    objs = []
    procs = []
    os.makedirs(os.path.join(_HERE, 'build'), exist_ok=True)
    for src in sources():
        obj = os.path.join(_HERE, 'build', os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd) + '\n' + out)
    tmp = LIB_PATH + '.tmp'
    cmd = [nvcc, '-shared', '-o', tmp] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed: ' + ' '.join(cmd) + '\n' + r.stdout)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_l = ctypes.c_int64
c_f = ctypes.c_float
c_d = ctypes.c_double
c_u64 = ctypes.c_uint64


class Seg(ctypes.Structure):
    """dyg_seg_t (include/dygb200.h)."""
    _fields_ = [('kind', ctypes.c_int32), ('width', ctypes.c_int32), ('group', ctypes.c_int32),
                ('ld', ctypes.c_int32), ('ld2', ctypes.c_int32), ('tq_div', ctypes.c_int32),
                ('ptr', c_p), ('idx', c_p), ('ptr2', c_p), ('idx2', c_p),
                ('dt', c_p), ('mask_ids', c_p), ('w', c_p), ('b', c_p), ('t_query', c_p)]


class ProjSide(ctypes.Structure):
    """dyg_proj_side_t (include/dygb200.h)."""
    _fields_ = [('ids', c_p), ('eids', c_p), ('t_nbr', c_p), ('cnt_a', c_p), ('cnt_b', c_p),
                ('tokens', ctypes.c_int64), ('ntok', ctypes.c_int32), ('tok_off', ctypes.c_int32)]


class Planes(ctypes.Structure):
    """dyg_planes_t (include/dygb200.h)."""
    _fields_ = [('hi', c_p), ('mid', c_p), ('ld', ctypes.c_int64)]


class TgnStep(ctypes.Structure):
    """dyg_tgn_step_t (include/dygb200.h), field for field."""
    _fields_ = ([('he', c_p), ('indptr', c_p), ('num_nodes', ctypes.c_int64), ('src', c_p), ('dst', c_p), ('t', c_p), ('eid', c_p),
                 ('neg', c_p), ('roots', c_p)] +
                [(n, ctypes.c_int32) for n in ('B', 'R', 'k', 'H', 'G', 'check_time')] +
                [('node_raw', c_p), ('ld_node', ctypes.c_int32), ('edge_raw', c_p), ('ld_edge', ctypes.c_int32),
                 ('F', ctypes.c_int32), ('E', ctypes.c_int32), ('T', ctypes.c_int32)] +
                [(n, c_p) for n in ('memory', 'last_update', 'mem_view', 'lu_view', 'pending', 'winner', 'msg_store', 'msg_time', 'flag',
                                    'time_w', 'time_b', 't0')] +
                [(n, Planes) for n in ('wqk', 'wvr', 'm1', 'm2', 'w_ih', 'w_hh', 'p1')] +
                [('cq', c_p), ('rbias', c_p), ('ln_g', c_p), ('ln_b', c_p), ('ln_eps', ctypes.c_float)] +
                [(n, c_p) for n in ('m1_b', 'm2_b', 'b_ih', 'b_hh', 'p1_b', 'p2_w', 'p2_b', 'pair_a', 'pair_b')] +
                [('P', ctypes.c_int32)] +
                [(n, c_p) for n in ('nbr_ids', 'nbr_eids', 'nbr_t', 'feat', 'qk', 'o', 'msg', 'hnew', 'ph')] +
                [(n, Planes) for n in ('feat_pl', 'msg_pl', 's_pl', 'y_pl', 'h1_pl')] +
                [(n, c_p) for n in ('emb', 'prob', 'barrier', 'phase_ns')])


# name -> argtypes, exactly the prototypes of include/dygb200.h (tests check the symbol list against the header)
SIGNATURES = {
    'dyg_csr_degrees': [c_p, c_p, c_l, c_l, c_p, c_p],
    'dyg_radix_digit_hist': [c_p, c_i, c_l, c_p, c_p],
    'dyg_radix_sort_pass': [c_p, c_p, c_p, c_p, c_i, c_l, c_i, c_p, c_p, c_p],
    'dyg_csr_pack': [c_p, c_p, c_p, c_p, c_p, c_l, c_p, c_p],
    'dyg_csr_tia_tables': [c_p, c_p, c_l, c_d, c_p, c_p, c_p],
    'dyg_csr_tia_cum': [c_p, c_p, c_l, c_p, c_p],
    'dyg_csr_fence_build': [c_p, c_l, c_p, c_p],
    'dyg_count_before': [c_p, c_p, c_l, c_p, c_l, c_p, c_p, c_l, c_p, c_p],
    'dyg_sample_recent': [c_p, c_p, c_l, c_p, c_l, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p, c_p],
    'dyg_sample_indexed': [c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p, c_p],
    'dyg_draw_uniform': [c_p, c_p, c_l, c_i, c_p, c_p],
    'dyg_draw_tia': [c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_p, c_p],
    'dyg_philox_uniform': [c_u64, c_u64, c_l, c_p, c_p],
    'dyg_sample_random': [c_p, c_p, c_l, c_p, c_l, c_p, c_p, c_p, c_p, c_l, c_i, c_u64, c_u64, c_p, c_p, c_p, c_p],
    'dyg_cum_fence_build': [c_p, c_l, c_p, c_p],
    'dyg_first_hop_pad': [c_p, c_p, c_l, c_p, c_l, c_p, c_p, c_l, c_i, c_i, c_p, c_p, c_p, c_p, c_p, c_i, c_p],
    'dyg_cooc_count': [c_p, c_i, c_p, c_i, c_l, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    'dyg_time_encode': [c_p, c_l, c_p, c_p, c_i, c_p, c_p],
    'dyg_time_encode_bwd': [c_p, c_l, c_p, c_p, c_i, c_p, c_l, c_p, c_p, c_p],
    'dyg_linear': [ctypes.POINTER(Seg), c_i, c_p, c_i, c_p, c_p, c_i, c_p, c_i, c_l, c_i, c_i, c_i, c_i, c_i, c_p],
    'dyg_linear_tc_tile': [c_i],
    'dyg_linear_tc': [ctypes.POINTER(Seg), c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_i, c_p, c_i, c_l, c_i, c_i, c_i, c_i, c_i, c_p],
    'dyg_gemm_bf16x3': [c_p, c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_l, c_i, c_i, c_i, c_p],
    'dyg_ln_ffn_bf16x3': [c_p, c_i, c_p, c_p, c_f, c_p, c_p, c_i, c_p, c_p, c_p, c_i, c_p, c_p, c_i, c_l, c_i, c_i, c_p, c_l, c_p],
    'dyg_ln_gemm_bf16x3': [c_p, c_i, c_p, c_p, c_f, c_p, c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_i, c_l, c_i, c_i, c_i, c_p, c_l, c_p],
    'dyg_split_bf16': [c_p, c_i, c_l, c_i, c_p, c_p, c_i, c_p],
    'dyg_layernorm_split': [c_p, c_i, c_p, c_p, c_f, c_p, c_i, c_p, c_p, c_i, c_l, c_i, c_p],
    'dyg_patch_project_stages': [c_i, c_i, c_i, c_i, c_i, c_p],
    'dyg_patch_project': [ctypes.POINTER(ProjSide), c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p,
                          c_i, c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_p, c_i, c_p],
    'dyg_layernorm': [c_p, c_i, c_p, c_i, c_i, c_p, c_p, c_p, c_f, c_p, c_i, c_l, c_i, c_p],
    'dyg_gather_rows': [c_p, c_i, c_p, c_i, c_p, c_l, c_i, c_p, c_i, c_p],
    'dyg_temporal_attend': [c_p, c_i, c_l, c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_i,
                            c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_i, c_p, c_i, c_p, c_p],
    'dyg_temporal_attend_bwd': [c_p, c_i, c_l, c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_i,
                                c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p],
    'dyg_seq_attention': [c_p, c_i, c_l, c_i, c_i, c_i, c_p, c_i, c_p],
    'dyg_seq_attention_tc': [c_p, c_i, c_l, c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_i, c_p],
    'dyg_seq_attention_fold': [c_p, c_p, c_i, c_i, c_i, c_i, c_l, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_p, c_i, c_p],
    'dyg_attn_block': [c_p, c_i, c_p, c_p, c_f, c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_p, c_l, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_l, c_p],
    'dyg_gemm_dw': [c_p, c_i, c_p, c_i, c_l, c_i, c_i, c_p, c_i, c_p, c_p],
    'dyg_gemm_dx': [c_p, c_i, c_p, c_i, c_p, c_i, c_l, c_i, c_i, c_p, c_i, c_p],
    'dyg_linear_bwd': [c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_l, c_i, c_i, c_p, c_i, c_p, c_i, c_p, c_p],
    'dyg_layernorm_bwd': [c_p, c_i, c_p, c_f, c_p, c_i, c_p, c_i, c_p, c_p, c_l, c_i, c_p],
    'dyg_gelu_fwd': [c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_l, c_i, c_p],
    'dyg_gelu_bwd': [c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_l, c_i, c_p],
    'dyg_seq_attention_train_fwd': [c_p, c_i, c_l, c_i, c_i, c_i, c_p, c_p, c_p, c_i, c_p],
    'dyg_seq_attention_train_bwd': [c_p, c_i, c_l, c_i, c_i, c_i, c_p, c_p, c_p, c_i, c_p, c_i, c_p],
    'dyg_mean_tokens': [c_p, c_l, c_i, c_i, c_i, c_i, c_p, c_i, c_p],
    'dyg_tgn_persist': [c_p, c_l, c_p, c_p, c_p, c_p, c_p, c_i, c_p],
    'dyg_tgn_select_last': [c_p, c_p, c_l, c_p, c_p],
    'dyg_tgn_build_messages': [c_p, c_p, c_p, c_p, c_l, c_p, c_p, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_p, c_i,
                               c_p, c_i, c_p],
    'dyg_tgn_cell_commit': [c_p, c_p, c_i, c_p, c_p, c_p, c_l, c_p, c_p, c_p, c_p, c_p, c_i, c_p, c_i, c_i,
                            c_p, c_p, c_p],
    'dyg_gru_update_fwd': [c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p, c_p, c_p, c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_l, c_p],
    'dyg_gru_update_bwd': [c_p, c_p, c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_l, c_i, c_p],
    'dyg_tgn_step': [ctypes.POINTER(TgnStep), c_p],
    'dyg_tgn_check_time': [c_p, c_l, c_p, c_p, c_p, c_p, c_p],
    'dyg_jodie_project': [c_p, c_i, c_p, c_p, c_p, c_l, c_i, c_f, c_f, c_p, c_p, c_p, c_i, c_p],
}


def load():
    """Return the ctypes handle, building the library first if it is missing or stale."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if _stale():
            # no silent fallback, and no stale binary either: SIGNATURES below describe the CURRENT header, so binding them
            # to an older build would pass mis-typed arguments to the kernels
            try:
                build()
            except Exception as e:
                what = 'is out of date' if os.path.exists(LIB_PATH) else 'is missing'
                raise RuntimeError(f'libdygb200.so {what} and could not be rebuilt: {e}') from e
        lib = ctypes.CDLL(LIB_PATH)
        lib.dyg_last_error.restype = ctypes.c_char_p
        lib.dyg_last_error.argtypes = []
        lib.dyg_abi_version.restype = c_i
        lib.dyg_abi_version.argtypes = []
        lib.dyg_ln_ffn_workspace_bytes.restype = c_l
        lib.dyg_csr_fence_entries.restype = c_l
        lib.dyg_csr_fence_entries.argtypes = [c_l]
        lib.dyg_ln_ffn_workspace_bytes.argtypes = []
        lib.dyg_attn_block_workspace_bytes.restype = c_l
        lib.dyg_attn_block_workspace_bytes.argtypes = [c_i]
        lib.dyg_radix_sort_workspace_entries.restype = c_l
        lib.dyg_radix_sort_workspace_entries.argtypes = [c_l]
        lib.dyg_tgn_step_sizeof.restype = c_l
        lib.dyg_tgn_step_sizeof.argtypes = []
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = c_i
            fn.argtypes = args
        if lib.dyg_tgn_step_sizeof() != ctypes.sizeof(TgnStep):
            raise RuntimeError('dyg_tgn_step_t layout differs between include/dygb200.h and dyglib_b200/_native.py')
        if lib.dyg_abi_version() != ABI_VERSION:
            raise RuntimeError(f'libdygb200.so ABI version {lib.dyg_abi_version()} != {ABI_VERSION} (include/dygb200.h)')
        _lib = lib
        return lib


def check(rc: int):
    if rc != 0:
        msg = load().dyg_last_error().decode()
        if rc == 2:
            raise ValueError(msg)
        raise RuntimeError(msg)
