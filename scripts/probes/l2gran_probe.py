"""Probe: does cudaLimitMaxL2FetchGranularity change the sampler sweep / headline?  python scripts/l2gran_probe.py <bytes> <bench args...>"""
import ctypes, sys, subprocess, os
gran = int(sys.argv[1])
import torch
torch.cuda.init()
rt = ctypes.CDLL('libcudart.so')
val = ctypes.c_size_t()
rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
print('before', val.value, flush=True)
rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran))
rt.cudaDeviceGetLimit(ctypes.byref(val), 5)
print('rc', rc, 'after', val.value, flush=True)
sys.argv = ['bench.py'] + sys.argv[2:]
sys.path.insert(0, '.')
import runpy
runpy.run_path('bench.py', run_name='__main__')
