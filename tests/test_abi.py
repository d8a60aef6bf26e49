"""CPU tests: the C-ABI library builds, loads, and exports exactly the symbols include/gglb200.h declares.
No compute calls (there is no GPU here); argument validation paths that return before any launch are exercised."""
import ctypes
import os
import re

from gammagl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'gglb200.h')).read()
    return sorted(set(re.findall(r'\b(gglb200_[a-z0-9_]+)\s*\(', hdr)))


def test_header_symbols_all_exported_and_bound():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(lib, s), 'libgglb200.so does not export %s' % s
        assert s in _lib.SIGNATURES, 'ctypes binding lacks %s' % s
    assert sorted(_lib.SIGNATURES) == syms


def test_exports_nothing_else():
    import subprocess
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.lib_path()], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if ' T ' in l)
    assert exported == declared_symbols()


def test_version_and_sizes():
    lib = _lib.load()
    assert lib.gglb200_abi_version() == 2
    assert lib.gglb200_csr_plan_ints(0) >= 4
    assert lib.gglb200_csr_build_workspace_bytes(1000, 10) >= 3 * 4000
    assert lib.gglb200_spmm_workspace_bytes(100, 16, 0) == 0          # no row can be long
    assert lib.gglb200_spmm_workspace_bytes(1 << 20, 16, 2) > lib.gglb200_spmm_workspace_bytes(1 << 20, 16, 0)


def test_argument_validation_without_gpu():
    lib = _lib.load()
    st = lib.gglb200_spmm_csr(7, None, None, None, None, None, 1, None, None, None, 0, None, 4, 4, 4, 4, 0, None, 0, None)
    assert st == -1 and b'unknown reduce' in lib.gglb200_last_error()
    st = lib.gglb200_csr_build(None, None, -1, 4, 4, None, None, None, None, None, 0, None)
    assert st == -1
    st = lib.gglb200_csr_build(None, None, 1 << 31, 4, 4, None, None, None, None, None, 0, None)
    assert st == -4
    st = lib.gglb200_spmm_csr(0, None, None, None, None, None, 1, None, None, None, 0, None, 4, 4, 4, 4, 99, None, 0, None)
    assert st == -1  # NULL rowptr is caught before the dtype switch
    try:
        _lib.check(st, 'spmm')
    except _lib.GglError as e:
        assert 'status -1' in str(e)
    else:
        raise AssertionError('check() must raise')


def test_cpu_tensors_are_rejected_loudly():
    import pytest
    import torch
    from gammagl_b200 import mpops
    x = torch.ones(3, 2)
    ids = torch.tensor([0, 0, 1])
    with pytest.raises(_lib.GglError):
        mpops.unsorted_segment_sum(x, ids, 2)
    with pytest.raises(_lib.GglError):
        mpops.gspmm(torch.tensor([[0, 1], [1, 0]]), None, torch.ones(2, 4))


def test_new_entry_points_validate_without_gpu():
    """gglb200_sort_edges / unique_edges / segment_expand_bwd: size queries and argument errors return before any launch."""
    lib = _lib.load()
    assert lib.gglb200_sort_edges_workspace_bytes(1000, 50, 50) >= 2 * 8000 + 2 * 4000
    assert lib.gglb200_sort_edges_workspace_bytes(0, 0, 0) > 0
    st = lib.gglb200_sort_edges(None, None, -1, 4, 4, None, None, None, None, None, None, None, 0, None)
    assert st == -1 and b'negative' in lib.gglb200_last_error()
    st = lib.gglb200_sort_edges(None, None, 5, 4, 4, None, None, None, None, None, None, None, 0, None)
    assert st == -1 and b'must not be NULL' in lib.gglb200_last_error()
    st = lib.gglb200_sort_edges(None, None, 1 << 31, 4, 4, None, None, None, None, None, None, None, 0, None)
    assert st == -4
    assert lib.gglb200_unique_edges(None, None, None, 0, None, None, None) == 0          # nothing to do
    assert lib.gglb200_unique_edges(None, None, None, 3, None, None, None) == -1
    assert lib.gglb200_segment_expand_bwd(2, None, None, None, None, None, 4, 4, 4, 0, None) == -1   # MAX is not a gather
    assert lib.gglb200_segment_expand_bwd(0, None, None, None, None, None, 0, 4, 4, 0, None) == 0    # E == 0
    assert lib.gglb200_segment_expand_bwd(0, None, None, None, None, None, 4, 4, 4, 0, None) == -1   # NULL pointers
    assert lib.gglb200_segment_expand_bwd(0, None, None, None, None, None, 4, 4, 4, 99, None) == -2  # unknown dtype (GGL_EUNSUPPORTED)


def test_round2_entry_points_validate_without_gpu():
    """tcgen05 Linear / sort-free reduce / sampler: size queries are host arithmetic, argument errors return before any launch."""
    lib = _lib.load()
    # K = 100 -> 25 sixteen-byte chunks -> 2 passes of 14 (or 4 passes of 8 in the 32-column plan: the larger of the two is reserved);
    # N = 256 -> 4 chunks of 64 columns; [W_hi | W_lo] tile pairs + alignment slack
    assert lib.gglb200_linear_workspace_bytes(100, 256) == max(2 * 4 * 2 * 14 * 64 * 16, 4 * 4 * 2 * 8 * 64 * 16) + 256
    assert lib.gglb200_linear_workspace_bytes(8, 16) == 1 * 1 * 2 * 4 * 64 * 16 + 256
    assert lib.gglb200_linear_workspace_bytes(0, 16) == 0
    assert lib.gglb200_linear_tf32x3(None, None, None, 0, None, 0, 8, 8, None, 0, None) == 0        # M == 0: nothing to do
    assert lib.gglb200_linear_tf32x3(None, None, None, 0, None, 4, 8, 8, None, 0, None) == -1       # NULL pointers
    assert b'NULL' in lib.gglb200_last_error()
    assert lib.gglb200_linear_tf32x3(None, None, None, 0, None, 4, 6, 8, None, 0, None) == -1       # checked before the shape: pointers
    assert lib.gglb200_segment_reduce_workspace_bytes(_lib.GGL_MAX, 10, 4) == 10 * 4 * 8 + 256
    assert lib.gglb200_segment_reduce_workspace_bytes(_lib.GGL_SUM, 10, 64) == 0                    # wide sums need no replicas
    assert lib.gglb200_segment_reduce_workspace_bytes(_lib.GGL_SUM, 10, 1) == 10 * 4 * 8 + 256       # 8 replicas of narrow sums


def test_utils_reject_cpu_tensors():
    import pytest
    import torch
    from gammagl_b200 import layers, utils
    ei = torch.tensor([[0, 1], [1, 0]])
    with pytest.raises(_lib.GglError):
        utils.sort_edge_index(ei)
    with pytest.raises(_lib.GglError):
        utils.coalesce(ei)
    with pytest.raises(RuntimeError):
        layers.degree(ei[0], 2)
