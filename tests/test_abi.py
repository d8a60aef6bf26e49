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
    assert lib.gglb200_abi_version() == 1
    assert lib.gglb200_csr_plan_ints(0) >= 4
    assert lib.gglb200_csr_build_workspace_bytes(1000, 10) >= 3 * 4000
    assert lib.gglb200_spmm_workspace_bytes(100, 16, 0) == 0          # no row can be long
    assert lib.gglb200_spmm_workspace_bytes(1 << 20, 16, 2) > lib.gglb200_spmm_workspace_bytes(1 << 20, 16, 0)


def test_argument_validation_without_gpu():
    lib = _lib.load()
    st = lib.gglb200_spmm_csr(7, None, None, None, None, None, 1, None, None, None, 0, None, 4, 4, 4, 4, 0, None, 0, None)
    assert st == -1 and b'unknown reduce' in lib.gglb200_last_error()
    st = lib.gglb200_csr_build(None, None, -1, 4, None, None, None, None, None, 0, None)
    assert st == -1
    st = lib.gglb200_csr_build(None, None, 1 << 31, 4, None, None, None, None, None, 0, None)
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
