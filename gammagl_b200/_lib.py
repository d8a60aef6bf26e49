"""ctypes binding of libgglb200.so (the C ABI declared in include/gglb200.h).

There is no CPU or pure-torch fallback anywhere in this package: if the shared library cannot be
built/loaded, importing the operators raises.  (The reference swallows extension errors and silently
falls back to scatter ops, /root/reference/gammagl/mpops/torch.py:83-86 -- deliberately not mirrored.)
"""
import ctypes
import os

from . import build as _build

GGL_SUM, GGL_MEAN, GGL_MAX = 0, 1, 2
GGL_F32, GGL_BF16, GGL_F16, GGL_F64, GGL_I8, GGL_I16, GGL_I32, GGL_I64, GGL_U8 = range(9)
LONG_ROW = 1024

_c = ctypes
_vp, _i64, _i32, _sz, _int = _c.c_void_p, _c.c_int64, _c.c_int32, _c.c_size_t, _c.c_int

# name -> (restype, argtypes); must list every symbol declared in include/gglb200.h
SIGNATURES = {
    'gglb200_abi_version': (_int, []),
    'gglb200_last_error': (_c.c_char_p, []),
    'gglb200_launch_count': (_i64, []),
    'gglb200_set_l2_fetch_granularity': (_int, [_int, _c.POINTER(_int)]),
    'gglb200_csr_build_workspace_bytes': (_sz, [_i64, _i64]),
    'gglb200_csr_plan_ints': (_sz, [_i64]),
    'gglb200_csr_build': (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gglb200_ind2ptr': (_int, [_vp, _i64, _i64, _vp, _vp]),
    'gglb200_ptr2ind': (_int, [_vp, _i64, _i64, _vp, _vp]),
    'gglb200_permute_rows': (_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    'gglb200_sort_edges_workspace_bytes': (_sz, [_i64, _i64, _i64]),
    'gglb200_sort_edges': (_int, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    'gglb200_unique_edges': (_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp]),
    'gglb200_spmm_workspace_bytes': (_sz, [_i64, _i64, _int]),
    'gglb200_spmm_workspace_bytes_for_chunks': (_sz, [_i64, _i64, _int]),
    'gglb200_spmm_csr': (_int, [_int, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i64, _int,
                                _vp, _sz, _vp]),
    'gglb200_spmm_csr_epilogue': (_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _int, _i64, _i64, _i64, _i64, _int, _vp, _sz, _vp]),
    'gglb200_spmm_csr_peer': (_int, [_int, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _i64, _i64, _i64, _int, _vp, _sz, _vp]),
    'gglb200_spmm_csr_max_edge': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _int, _vp, _sz, _vp]),
    'gglb200_spmm_max_edge_bwd': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    'gglb200_segment_reduce_workspace_bytes': (_sz, [_int, _i64, _i64]),
    'gglb200_segment_reduce': (_int, [_int, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64, _i64, _int, _vp, _sz, _vp]),
    'gglb200_linear_workspace_bytes': (_sz, [_i64, _i64]),
    'gglb200_linear_tf32x3': (_int, [_vp, _vp, _vp, _int, _vp, _i64, _i64, _i64, _vp, _sz, _vp]),
    'gglb200_sample_workspace_bytes': (_sz, [_i64]),
    'gglb200_sample_count': (_int, [_vp, _vp, _i64, _i64, _i64, _int, _vp, _vp, _vp, _sz, _vp]),
    'gglb200_sample_pick': (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _c.c_uint64, _vp, _vp, _vp, _vp]),
    'gglb200_gather_rows_peer': (_int, [_vp, _i32, _vp, _vp, _i64, _i64, _vp]),
    'gglb200_scale_rows': (_int, [_vp, _vp, _vp, _i64, _i64, _int, _vp]),
    'gglb200_degree_power': (_int, [_vp, _vp, _i64, _c.c_float, _vp]),
    'gglb200_segment_gather_bwd': (_int, [_int, _vp, _vp, _vp, _vp, _i64, _i64, _int, _vp]),
    'gglb200_segment_expand_bwd': (_int, [_int, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    'gglb200_segment_max_bwd': (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i64, _int, _vp]),
    'gglb200_scale_rows_by_count': (_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    'gglb200_spmm_max_bwd': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp]),
    'gglb200_bspmm_bwd_w': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    'gglb200_gat_workspace_bytes': (_sz, [_i64, _i32, _i32]),
    'gglb200_gat_fwd': (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _c.c_float, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _sz, _vp]),
    'gglb200_gat_bwd_src': (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _c.c_float, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _sz, _vp]),
    'gglb200_gat_bwd': (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _c.c_float, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _sz, _vp]),
    'gglb200_gat_bwd_dst': (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _c.c_float, _vp, _i64, _i64, _i32, _i32, _vp, _sz, _vp]),
    'gglb200_edge_softmax_stats_oneshot': (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp]),
    'gglb200_edge_softmax_dot_oneshot': (_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp]),
    'gglb200_edge_softmax_workspace_bytes': (_sz, [_i64, _i32, _i64]),
    'gglb200_edge_softmax_fwd': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp, _sz, _vp]),
    'gglb200_edge_softmax_bwd': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _vp, _sz, _vp]),
}

_lib = None


class GglError(RuntimeError):
    pass


def lib_path():
    return _build.LIB


def load():
    """Build (if stale and nvcc is available) and dlopen libgglb200.so.  Raises on any failure."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()
    if not os.path.exists(path):
        raise GglError('libgglb200.so is missing (%s) and could not be built' % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.gglb200_abi_version() != 2:
        raise GglError('libgglb200.so ABI version mismatch')
    _lib = lib
    return lib


def check(status, what=''):
    if status != 0:
        msg = load().gglb200_last_error().decode('utf-8', 'replace')
        raise GglError('libgglb200 %s failed (status %d): %s' % (what, status, msg))


_tuned_devices = set()


_previous_l2_fetch = {}


def tune_device(device_index):
    """Once per device: 32-byte L2 fetch granularity for the random row gathers (env GGLB200_L2_FETCH=32|64|128,
    0 leaves the driver default untouched).  Must be called with `device_index` current.
    This is a device-wide limit, so it also applies to the host application's other kernels; measured on B200 at the C3 shape
    (tools/bench_layer.py, profiles/bench_layer_r2.json): the fp32 cuBLAS GEMM next to the aggregation runs 2.95 / 2.97 / 2.97 ms at
    32 / 64 / 128 bytes -- no penalty --, the aggregation 5.38 / 5.46 / 5.49 ms.  restore_device() puts the previous value back."""
    if device_index in _tuned_devices:
        return
    _tuned_devices.add(device_index)
    want = int(os.environ.get('GGLB200_L2_FETCH', '32'))
    if want:
        prev = ctypes.c_int(0)
        check(load().gglb200_set_l2_fetch_granularity(want, ctypes.byref(prev)), 'set_l2_fetch_granularity')
        _previous_l2_fetch[device_index] = prev.value


def restore_device(device_index):
    """Undo tune_device() on the current device (the L2 fetch granularity the process found)."""
    prev = _previous_l2_fetch.pop(device_index, None)
    _tuned_devices.discard(device_index)
    if prev in (32, 64, 128):
        check(load().gglb200_set_l2_fetch_granularity(prev, None), 'set_l2_fetch_granularity')


def launch_count():
    return int(load().gglb200_launch_count())
