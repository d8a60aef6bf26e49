"""CPU test: the binding stub printed in INTEGRATION.md (section B) stays consistent with the library -- it compiles, the
argument lists it declares have the arity of the ctypes table (hence of include/gglb200.h), and every gglb200_* name it or
the mapping table mentions is an exported entry point."""
import os
import re

from gammagl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOC = open(os.path.join(ROOT, 'INTEGRATION.md')).read()


def python_blocks():
    return re.findall(r'```python\n(.*?)```', DOC, flags=re.S)


def test_stub_compiles_and_matches_signatures():
    stub = [b for b in python_blocks() if 'torch_ext_b200' in b]
    assert len(stub) == 1
    compile(stub[0], 'INTEGRATION.md', 'exec')
    for name, args in re.findall(r'_lib\.(gglb200_\w+)\.argtypes = \[(.*?)\]', stub[0], flags=re.S):
        n = len([a for a in args.replace('\n', ' ').split(',') if a.strip()])
        assert name in _lib.SIGNATURES, name
        assert n == len(_lib.SIGNATURES[name][1]), (name, n, len(_lib.SIGNATURES[name][1]))


def test_every_mentioned_entry_point_exists():
    lib = _lib.load()
    mentioned = set(re.findall(r'\b(gglb200_[a-z0-9_]+)\b', DOC))
    # family shorthands used in the mapping table
    shorthand = {'gglb200_edge_softmax_fwd', 'gglb200_spmm_workspace_bytes'}
    for name in sorted(mentioned):
        if name.endswith('_') or name in ('gglb200_b200',):
            continue
        assert hasattr(lib, name) or any(name.startswith(s) for s in shorthand), name
    header = open(os.path.join(ROOT, 'include', 'gglb200.h')).read()
    declared = set(re.findall(r'\b(gglb200_[a-z0-9_]+)\s*\(', header))
    n_doc = int(re.search(r'\((\d+) `extern "C"` entry points', DOC).group(1))
    assert n_doc == len(declared)
