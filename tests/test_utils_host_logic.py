"""CPU test of the HOST logic of gammagl_b200.utils (sort_edge_index / coalesce / to_undirected / is_undirected): the three
places where it touches the GPU -- sort_pairs, permute_rows, unique_pairs -- and the segment sum are replaced by torch CPU
stand-ins with the same contracts (include/gglb200.h), and the composed functions must reproduce the reference's own KATs
and the outputs of the reference functions (tests/golden/utils_*.npz).  The kernels themselves are covered by -m gpu."""
import glob
import os

import numpy as np
import pytest
import torch

from gammagl_b200 import mpops, utils

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RAND = sorted(glob.glob(os.path.join(GOLDEN, 'utils_rand_*.npz')))


@pytest.fixture()
def U(monkeypatch):
    def sort_pairs(major, minor, n_major, n_minor, want_sorted=True, want_groups=False):
        key = major.to(torch.int64) * int(n_minor) + minor.to(torch.int64)
        perm = torch.from_numpy(np.argsort(key.numpy(), kind='stable'))
        sm, sn = major[perm], minor[perm]
        k = key[perm]
        first = torch.ones_like(k, dtype=torch.bool)
        first[1:] = k[1:] != k[:-1]
        group = torch.cumsum(first.to(torch.int64), 0) - 1
        n_unique = torch.tensor([int(first.sum())])
        return (sm if want_sorted else None, sn if want_sorted else None, perm.to(torch.int32),
                group if want_groups else None, n_unique if want_groups else None)

    def unique_pairs(sm, sn, group, n_unique):
        um, un = torch.empty(n_unique, dtype=torch.int64), torch.empty(n_unique, dtype=torch.int64)
        first = torch.ones_like(group, dtype=torch.bool)
        first[1:] = group[1:] != group[:-1]
        um[group[first]], un[group[first]] = sm[first], sn[first]
        return um, un

    def segment_sum(x, ids, n):
        return torch.zeros((n,) + tuple(x.shape[1:]), dtype=x.dtype).index_add_(0, ids, x)
    monkeypatch.setattr(utils, 'require_cuda', lambda t, name: None)
    monkeypatch.setattr(utils, 'sort_pairs', sort_pairs)
    monkeypatch.setattr(utils, 'permute_rows', lambda t, perm: t[perm.to(torch.int64)])
    monkeypatch.setattr(utils, 'unique_pairs', unique_pairs)
    monkeypatch.setattr(mpops, 'unsorted_segment_sum', segment_sum)
    return utils


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_reference_kats(U):
    z = np.load(os.path.join(GOLDEN, 'utils_kat.npz'))
    assert U.sort_edge_index(t(z['sort_ei'])).tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]               # test_sort_edge_index.py:10-15
    out, attr = U.sort_edge_index(t(z['sort_ei']), t(z['sort_attr']))
    assert attr.tolist() == [[4], [3], [2], [1]]
    assert U.coalesce(t(z['co_ei'])).tolist() == [[0, 1], [1, 0]]                                      # test_coalesce.py:13-21
    ei, attr = U.coalesce(t(z['co_ei']), t(z['co_attr']))
    assert ei.tolist() == [[0, 1], [1, 0]] and attr.tolist() == [[-4.0, 4.0], [2.0, 5.0]]
    row, col = torch.tensor([0, 1, 0]), torch.tensor([1, 0, 0])                                         # test_undirected.py:5-36
    assert U.is_undirected(torch.stack([row, col])) and U.is_undirected(torch.stack([row, col]), torch.tensor([0, 0, 1]))
    assert not U.is_undirected(torch.stack([row, col]), torch.tensor([0, 1, 1]))
    assert not U.is_undirected(t(z['und_ei']))
    ei, attr = U.to_undirected(t(z['und_ei']), t(z['und_attr']))
    assert ei.tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]] and attr.tolist() == [[3., 3.], [3., 3.], [-5., 4.], [-5., 4.]]
    assert U.to_undirected(t(z['und_ei']), 3).tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]
    ei, attrs = U.coalesce(t(z['co_ei']), [t(z['co_attr']), t(z['co_attr'][:, 0].copy())])
    assert attrs[0].tolist() == [[-4.0, 4.0], [2.0, 5.0]] and attrs[1].tolist() == [-4.0, 2.0]


@pytest.mark.parametrize('path', RAND, ids=[os.path.basename(p)[:-4] for p in RAND])
def test_reference_outputs(U, path):
    z = np.load(path)
    N = int(z['N'])
    for sbr, tag in ((True, 'row'), (False, 'col')):
        out, attr = U.sort_edge_index(t(z['ei']), t(z['attr']), N, sort_by_row=sbr)
        assert np.array_equal(out.numpy(), z['sort_%s_ei' % tag])
        if 'sort_%s_attr' % tag in z.files:
            assert np.array_equal(attr.numpy(), z['sort_%s_attr' % tag])
        ei, attr = U.coalesce(t(z['ei']), t(z['attr']), N, sort_by_row=sbr)
        assert np.array_equal(ei.numpy(), z['co_%s_ei' % tag]) and np.array_equal(attr.numpy(), z['co_%s_attr' % tag])
    ei, attr = U.to_undirected(t(z['ei']), t(z['attr']), N)
    assert np.array_equal(ei.numpy(), z['und_ei']) and np.array_equal(attr.numpy(), z['und_attr'])
    got = [U.is_undirected(t(z['ei']), None, N), U.is_undirected(ei, None, N), U.is_undirected(ei, attr, N)]
    assert got == z['is_und'].tolist()


def test_empty_and_single_edge(U):
    empty = torch.zeros((2, 0), dtype=torch.int64)
    assert U.sort_edge_index(empty, num_nodes=4).shape == (2, 0) and U.coalesce(empty, num_nodes=4).shape == (2, 0)
    assert U.coalesce(torch.tensor([[3], [1]])).tolist() == [[3], [1]]
