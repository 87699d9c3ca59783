"""ctypes loader of the plain-C MSDA oracle (oracle/msda_core.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, 'msda_core.c')
_LIB = os.path.join(_HERE, '_build', 'libmsda_core.so')
_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-o', _LIB, _SRC, '-lm'])
    return _LIB


def msda_forward(value, spatial_shapes, sampling_loc, attn_weight):
    """numpy fp32 in, numpy fp32 out; same argument meaning as oracle.mmcv_semantics.msda_core."""
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ub_oracle_msda_forward.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 7
        _lib.ub_oracle_msda_forward.restype = None
    value = np.ascontiguousarray(value, dtype=np.float32)
    loc = np.ascontiguousarray(sampling_loc, dtype=np.float32)
    w = np.ascontiguousarray(attn_weight, dtype=np.float32)
    shapes = np.ascontiguousarray(spatial_shapes, dtype=np.int64).reshape(-1, 2)
    start = np.concatenate(([0], np.cumsum(shapes[:, 0] * shapes[:, 1])[:-1])).astype(np.int64)
    B, Nv, H, D = value.shape
    _, Nq, _, L, P, _ = loc.shape
    out = np.empty((B, Nq, H * D), dtype=np.float32)
    _lib.ub_oracle_msda_forward(value.ctypes.data, shapes.ctypes.data, start.ctypes.data, loc.ctypes.data,
                                w.ctypes.data, out.ctypes.data, B, Nv, H, D, Nq, L, P)
    return out
