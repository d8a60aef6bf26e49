"""GPU parity tests (-m gpu) of the edge-list utilities in front of the path:
gammagl_b200.utils.sort_edge_index / coalesce (C ABI: gglb200_sort_edges, gglb200_unique_edges, gglb200_permute_rows)
against the reference's known-answer tests, outputs of the reference functions (tests/golden/utils_*.npz) and the numpy
oracle on seeded inputs.  Integer results bit-exact; merged attributes bit-exact too (same serial order as the oracle)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RAND = sorted(glob.glob(os.path.join(GOLDEN, 'utils_rand_*.npz')))
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def U():
    from gammagl_b200 import _lib, utils
    _lib.load()
    return utils


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_kat(U):
    z = np.load(os.path.join(GOLDEN, 'utils_kat.npz'))
    out = U.sort_edge_index(cu(z['sort_ei']))
    assert out.dtype == torch.int64 and out.cpu().tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]
    out, attr = U.sort_edge_index(cu(z['sort_ei']), cu(z['sort_attr']))
    assert out.cpu().tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]] and attr.cpu().tolist() == [[4], [3], [2], [1]]
    ei = U.coalesce(cu(z['co_ei']))
    assert ei.cpu().tolist() == [[0, 1], [1, 0]]
    ei, attr = U.coalesce(cu(z['co_ei']), cu(z['co_attr']))
    assert ei.cpu().tolist() == [[0, 1], [1, 0]] and attr.cpu().tolist() == [[-4.0, 4.0], [2.0, 5.0]]
    # list attributes: every entry is permuted / merged
    ei, attrs = U.coalesce(cu(z['co_ei']), [cu(z['co_attr']), cu(z['co_attr'][:, 0].copy())])
    assert attrs[0].cpu().tolist() == [[-4.0, 4.0], [2.0, 5.0]] and attrs[1].cpu().tolist() == [-4.0, 2.0]


@pytest.mark.parametrize('path', RAND, ids=[os.path.basename(p)[:-4] for p in RAND])
def test_reference_outputs(U, path):
    z = np.load(path)
    N = int(z['N'])
    for sbr, tag in ((True, 'row'), (False, 'col')):
        out, attr = U.sort_edge_index(cu(z['ei']), cu(z['attr']), N, sort_by_row=sbr)
        assert np.array_equal(out.cpu().numpy(), z['sort_%s_ei' % tag])
        if 'sort_%s_attr' % tag in z.files:
            assert np.array_equal(attr.cpu().numpy(), z['sort_%s_attr' % tag])
        ei, attr = U.coalesce(cu(z['ei']), cu(z['attr']), N, sort_by_row=sbr)
        assert np.array_equal(ei.cpu().numpy(), z['co_%s_ei' % tag])
        assert np.array_equal(attr.cpu().numpy(), z['co_%s_attr' % tag])


@pytest.mark.parametrize('N,E', [(1, 50), (5000, 200000), (1 << 20, 300000), (40, 100000)])
def test_vs_oracle_random(U, N, E):
    g = torch.Generator().manual_seed(N + E)
    ei = torch.randint(0, N, (2, E), generator=g)
    attr = torch.randn(E, 5, generator=g)
    iattr = torch.randint(-100, 100, (E,), generator=g)
    for sbr in (True, False):
        ref, perm = O.sort_edge_index(ei.numpy(), N, sbr)
        out, (a1, a2) = U.sort_edge_index(ei.to(DEV), [attr.to(DEV), iattr.to(DEV)], N, sort_by_row=sbr)
        assert np.array_equal(out.cpu().numpy(), ref)
        assert np.array_equal(a1.cpu().numpy(), attr.numpy()[perm])          # stable: the same order as the oracle's
        assert np.array_equal(a2.cpu().numpy(), iattr.numpy()[perm])
        rei, rattr = O.coalesce(ei.numpy(), attr.numpy(), N, sbr)
        cei, cattr = U.coalesce(ei.to(DEV), attr.to(DEV), N, sort_by_row=sbr)
        assert np.array_equal(cei.cpu().numpy(), rei)
        deg = np.bincount(np.unique(ei.numpy()[0] * N + ei.numpy()[1], return_inverse=True)[1])
        if deg.max() <= 1024:
            assert np.array_equal(cattr.cpu().numpy(), rattr)                # same serial order and rounding
        else:
            assert np.allclose(cattr.cpu().numpy(), rattr, rtol=1e-5, atol=1e-4)
        _, riattr = O.coalesce(ei.numpy(), iattr.numpy(), N, sbr)
        _, ciattr = U.coalesce(ei.to(DEV), iattr.to(DEV), N, sort_by_row=sbr)
        assert np.array_equal(ciattr.cpu().numpy(), riattr)


def test_edge_cases(U):
    empty = torch.zeros((2, 0), dtype=torch.int64, device=DEV)
    assert U.sort_edge_index(empty, num_nodes=4).shape == (2, 0)
    assert U.coalesce(empty, num_nodes=4).shape == (2, 0)
    one = torch.tensor([[3], [1]], device=DEV)
    assert U.coalesce(one).cpu().tolist() == [[3], [1]]
    same = torch.tensor([[2] * 3000, [2] * 3000], device=DEV)                  # one group longer than a split row
    ei, attr = U.coalesce(same, torch.ones(3000, device=DEV))
    assert ei.cpu().tolist() == [[2], [2]] and attr.cpu().tolist() == [3000.0]
    with pytest.raises(Exception):
        U.sort_edge_index(torch.zeros((2, 3), dtype=torch.int64))              # CPU tensor: no fallback


def test_sorted_input_is_unchanged_and_idempotent(U):
    g = torch.Generator().manual_seed(4)
    ei = torch.randint(0, 1000, (2, 50000), generator=g).to(DEV)
    a = U.sort_edge_index(ei, num_nodes=1000)
    assert torch.equal(U.sort_edge_index(a, num_nodes=1000), a)
    c = U.coalesce(ei, num_nodes=1000)
    assert torch.equal(U.coalesce(c, num_nodes=1000, is_sorted=True), c)
    key = c[0] * 1000 + c[1]
    assert bool((key[1:] > key[:-1]).all())


def test_undirected_kat_and_reference_outputs(U):
    """to_undirected / is_undirected (utils/undirected.py) against tests/utils/test_undirected.py:5-36 and reference outputs."""
    z = np.load(os.path.join(GOLDEN, 'utils_kat.npz'))
    row, col = torch.tensor([0, 1, 0], device=DEV), torch.tensor([1, 0, 0], device=DEV)
    assert U.is_undirected(torch.stack([row, col]))
    assert U.is_undirected(torch.stack([row, col]), torch.tensor([0, 0, 1], device=DEV))
    assert not U.is_undirected(torch.stack([row, col]), torch.tensor([0, 1, 1], device=DEV))
    assert not U.is_undirected(cu(z['und_ei']))
    ei, attr = U.to_undirected(cu(z['und_ei']), cu(z['und_attr']))
    assert ei.cpu().tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]
    assert attr.cpu().tolist() == [[3., 3.], [3., 3.], [-5., 4.], [-5., 4.]]
    assert U.to_undirected(cu(z['und_ei']), 3).cpu().tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]       # (edge_index, num_nodes) form
    for path in RAND:
        r = np.load(path)
        N = int(r['N'])
        ei, attr = U.to_undirected(cu(r['ei']), cu(r['attr']), N)
        assert np.array_equal(ei.cpu().numpy(), r['und_ei']) and np.array_equal(attr.cpu().numpy(), r['und_attr'])
        got = [U.is_undirected(cu(r['ei']), None, N), U.is_undirected(ei, None, N), U.is_undirected(ei, attr, N)]
        assert got == r['is_und'].tolist()
