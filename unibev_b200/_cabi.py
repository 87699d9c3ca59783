"""ctypes binding of libunibev_b200.so (the C ABI declared in include/unibev_b200.h).

There is no CPU fallback: if the shared library is missing, loading raises and every
op in this package fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libunibev_b200.so')

_p = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_f = ctypes.c_float

# name -> argtypes; every entry point of include/unibev_b200.h (tests check the header against this table)
PROTOTYPES = {
    'ub_version': ([], _i),
    'ub_last_error': ([], ctypes.c_char_p),
    'ub_launch_count': ([], _i64),
    'ub_launch_count_reset': ([], None),
    'ub_unsupported_count': ([], _i64),
    'ub_msda_fwd': ([_p] * 6 + [_i] * 7 + [_p], _i),
    'ub_msda_bwd': ([_p] * 9 + [_i] * 7 + [_p], _i),
    'ub_project_points': ([_p, _p, _p, _f, _f, _p, _p] + [_i] * 5 + [_p], _i),
    'ub_bev_sample_fwd': ([_p] * 3 + [_i] * 11 + [_p], _i),
    'ub_img_sample_fwd': ([_p] * 5 + [_i] * 13 + [_p], _i),
    'ub_bev_sample_bwd': ([_p] * 5 + [_i] * 11 + [_p], _i),
    'ub_img_sample_bwd': ([_p] * 7 + [_i] * 13 + [_p], _i),
    'ub_value_to_half': ([_p, _p] + [_i] * 4 + [_p], _i),
    'ub_bev_sample_win_fwd': ([_p] * 3 + [_i] * 12 + [_p, _i, _p], _i),
    'ub_bev_sample_win32_fwd': ([_p] * 3 + [_i] * 11 + [_p, _p], _i),
    'ub_hit_order': ([_p] * 8 + [_i] * 4 + [_p], _i),
    'ub_img_sample_win32_fwd': ([_p] * 7 + [_i] * 13 + [_p], _i),
    'ub_build_hits': ([_p] * 5 + [_i] * 3 + [_p], _i),
    'ub_img_sample_win_fwd': ([_p] * 8 + [_i] * 14 + [_p], _i),
    'ub_linear_tf32': ([_p] * 4 + [_i, _p, _p, _f, _p, _i, _p] + [_i] * 5 + [_p], _i),
    'ub_linear_tf32_dual': ([_p] * 4 + [_i, _p, _p, _f, _p, _i, _p] + [_i] * 5 + [_p], _i),
    'ub_linear_tf32x3': ([_p] * 5 + [_i, _p, _p, _f, _p, _i, _p] + [_i] * 5 + [_p], _i),
    'ub_linear_tf32x3_scatter': ([_p] * 5 + [_i, _p] + [_i] * 6 + [_p], _i),
    'ub_linear_f16x3': ([_p, _f] + [_p] * 5 + [_i, _p, _p, _f, _p, _i, _p, _i, _p] + [_i] * 7 + [_p], _i),
    'ub_linear_f16x3_dyn': ([_p, _p, _f, _f] + [_p] * 5 + [_i, _p, _p, _f, _p, _i, _p] + [_i] * 5 + [_p], _i),
    'ub_split_f16': ([_p, _p, _p, _p, _i, _i, _f, _p], _i),
    'ub_linear_simt': ([_p] * 4 + [_i, _p] + [_i] * 5 + [_p], _i),
    'ub_split_tf32': ([_p, _p, _p, _i64, _p], _i),
    'ub_linear_f16': ([_p] * 4 + [_i, _p, _p, _f, _p, _i, _p, _i, _p] + [_i] * 5 + [_p], _i),
    'ub_add_layernorm': ([_p] * 6 + [_i64, _i, _f, _p], _i),
    'ub_colsum': ([_p, _p, _i64, _i, _p], _i),
    'ub_layernorm_bwd': ([_p] * 7 + [_i64, _i, _f, _p], _i),
    'ub_dropout_add_layernorm_fwd': ([_p] * 6 + [_i64, _i, _f, _f, _p, _i, _p], _i),
    'ub_dropout_add_layernorm_bwd': ([_p] * 9 + [_i64, _i, _f, _f, _p], _i),
    'ub_add_layernorm16': ([_p] * 7 + [_i64, _i, _f, _p], _i),
    'ub_cnw_fuse': ([_p] * 8 + [_i64, _i, _i, _i, _i, _i, _p], _i),
    'ub_flatten_feats': ([_p, _p, _i, _p, _p, _i, _i, _i, _p], _i),
    'ub_flatten_feats_max': ([_p, _p, _i, _p, _p, _p, _i, _i, _i, _p], _i),
    'ub_flatten_feats16': ([_p, _p, _i, _p, _p, _p, _i, _i, _i, _p], _i),
    'ub_broadcast_rows': ([_p, _i64, _i, _i, _p, _p, _p], _i),
    'ub_voxelize_workspace_bytes': ([_i, ctypes.POINTER(ctypes.c_size_t)], _i),
    'ub_hard_voxelize': ([_p, _i, _i, ctypes.POINTER(_f), ctypes.POINTER(_f), _i, _i, _p, _p, _p, _p, _p,
                          ctypes.c_size_t, _p], _i),
    'ub_voxel_mean': ([_p, _p, _i, _i, _i, _i, _p, _p], _i),
}
# not part of the public header: process-wide performance knobs for A/B runs (tools/, sweeps); results never depend on them
_PRIVATE = {
    'ub_set_tuning': ([_i] * 3, _i),
    'ub_set_window_halo': ([_i], _i),
    'ub_set_bev_win32_buffers': ([_i], _i),
    'ub_set_gemm_cluster': ([_i], _i),
    'ub_set_gemm_x3': ([_i] * 4, _i),
    'ub_set_gemm_x3_pair': ([_i], _i),
    'ub_set_gemm_x3_prefetch': ([_i], _i),
    'ub_set_gemm_trace': ([_p], _i),
    'ub_set_pdl': ([_i], _i),
    'ub_set_img_two_windows': ([_i], _i),
    'ub_set_img_stage': ([_i], _i),
    'ub_set_img_vec_ref': ([_i], _i),
    'ub_set_gemm_stream_w_with_residual': ([_i], _i),
}

_lib = None


class UniBEVNativeError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UniBEVNativeError(
                f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a). unibev_b200 has no CPU or PyTorch fallback.')
        handle = ctypes.CDLL(LIB_PATH)
        for table in (PROTOTYPES, _PRIVATE):
            for name, (argtypes, restype) in table.items():
                fn = getattr(handle, name)
                fn.argtypes = argtypes
                fn.restype = restype
        _lib = handle
        # tuning knobs for A/B runs (defaults are the measured-best settings)
        for env, fn in (('UB_IMG_STAGE', 'ub_set_img_stage'), ('UB_IMG_TWO_WINDOWS', 'ub_set_img_two_windows'),
                        ('UB_IMG_VECREF', 'ub_set_img_vec_ref'), ('UB_PDL', 'ub_set_pdl')):
            if env in os.environ:
                getattr(handle, fn)(int(os.environ[env]))
        if 'UB_WIN32_BUFFERS' in os.environ:
            handle.ub_set_bev_win32_buffers(int(os.environ['UB_WIN32_BUFFERS']))
        if 'UB_X3_PREFETCH' in os.environ:
            handle.ub_set_gemm_x3_prefetch(int(os.environ['UB_X3_PREFETCH']))
        if 'UB_X3_PAIR' in os.environ:
            handle.ub_set_gemm_x3_pair(int(os.environ['UB_X3_PAIR']))
        if 'UB_X3' in os.environ:       # "inplace,direct,cluster,stagger_ns"
            handle.ub_set_gemm_x3(*[int(v) for v in os.environ['UB_X3'].split(',')])
    return _lib


UB_EUNSUPPORTED = -4
unsupported_log = []      # messages of the last UB_EUNSUPPORTED answers (diagnostics: which call fell back, and why)


class UnsupportedShape(UniBEVNativeError):
    """The specialised entry point has no kernel for this shape; callers switch to the generic one."""


def check(rc, what):
    if rc != 0:
        msg = lib().ub_last_error().decode('utf-8', 'replace')
        if rc == -1:
            raise ValueError(f'{what}: {msg}')
        if rc == UB_EUNSUPPORTED:
            unsupported_log.append(f'{what}: {msg}')
            del unsupported_log[:-32]
            raise UnsupportedShape(f'{what}: {msg}')
        raise UniBEVNativeError(f'{what} failed (code {rc}): {msg}')


def launch_count():
    return int(lib().ub_launch_count())


def reset_launch_count():
    lib().ub_launch_count_reset()


def unsupported_count():
    """Calls that returned UB_EUNSUPPORTED (and so sent their caller to a generic entry point) since the last reset."""
    return int(lib().ub_unsupported_count())


def set_tuning(which, tile_w=8, ctas_per_sm=0):
    """which: 0 = ub_bev_sample_fwd, 1 = ub_img_sample_fwd; tile = tile_w x 64/tile_w queries."""
    check(lib().ub_set_tuning(which, tile_w, ctas_per_sm), 'ub_set_tuning')
