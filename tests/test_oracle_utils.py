"""CPU tests (-m "not gpu"): pin oracle.sort_edge_index / oracle.coalesce (numpy restatements of
/root/reference/gammagl/utils/sort_edge_index.py:31-45 and coalesce.py:41-78) against the reference's own known-answer
tests and against outputs of the reference functions themselves (tests/golden/utils_*.npz, make_golden_utils.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RAND = sorted(glob.glob(os.path.join(GOLDEN, 'utils_rand_*.npz')))


def test_kat_sort_edge_index_and_coalesce():
    z = np.load(os.path.join(GOLDEN, 'utils_kat.npz'))
    out, perm = O.sort_edge_index(z['sort_ei'], 3)
    assert out.tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]                      # tests/utils/test_sort_edge_index.py:10-11
    assert z['sort_attr'][perm].tolist() == [[4], [3], [2], [1]]             # :15
    assert np.array_equal(out, z['sort_out'])
    ei, attr = O.coalesce(z['co_ei'], z['co_attr'], 2)
    assert ei.tolist() == [[0, 1], [1, 0]]                                   # tests/utils/test_coalesce.py:13-14, 19
    assert attr.tolist() == [[-4.0, 4.0], [2.0, 5.0]]                        # :20-21
    ei2, none = O.coalesce(z['co_ei'], None, 2)
    assert none is None and np.array_equal(ei2, ei)


@pytest.mark.parametrize('path', RAND, ids=[os.path.basename(p)[:-4] for p in RAND])
def test_oracle_matches_reference_outputs(path):
    z = np.load(path)
    N = int(z['N'])
    assert len(RAND) == 3
    for sbr, tag in ((True, 'row'), (False, 'col')):
        out, perm = O.sort_edge_index(z['ei'], N, sbr)
        assert np.array_equal(out, z['sort_%s_ei' % tag])
        if 'sort_%s_attr' % tag in z.files:                                  # no duplicate edges: order fully determined
            assert np.array_equal(z['attr'][perm], z['sort_%s_attr' % tag])
        ei, attr = O.coalesce(z['ei'], z['attr'], N, sbr)
        assert np.array_equal(ei, z['co_%s_ei' % tag])
        assert np.array_equal(attr, z['co_%s_attr' % tag])                   # integer-valued floats: exact in any order


def test_kat_undirected():
    """/root/reference/tests/utils/test_undirected.py:5-36."""
    z = np.load(os.path.join(GOLDEN, 'utils_kat.npz'))
    row, col = np.array([0, 1, 0]), np.array([1, 0, 0])
    assert O.is_undirected(np.stack([row, col]), None, 2)
    assert O.is_undirected(np.stack([row, col]), np.array([0, 0, 1]), 2)
    assert not O.is_undirected(np.stack([row, col]), np.array([0, 1, 1]), 2)
    assert not O.is_undirected(z['und_ei'], None, 3)
    ei, attr = O.to_undirected(z['und_ei'], z['und_attr'], 3)
    assert ei.tolist() == [[0, 1, 1, 2], [1, 0, 2, 1]]
    assert attr.tolist() == [[3., 3.], [3., 3.], [-5., 4.], [-5., 4.]]
    assert np.array_equal(ei, z['und_out']) and np.array_equal(attr, z['und_out_attr'])


@pytest.mark.parametrize('path', RAND, ids=[os.path.basename(p)[:-4] for p in RAND])
def test_undirected_oracle_matches_reference_outputs(path):
    z = np.load(path)
    N = int(z['N'])
    ei, attr = O.to_undirected(z['ei'], z['attr'], N)
    assert np.array_equal(ei, z['und_ei']) and np.array_equal(attr, z['und_attr'])
    got = [O.is_undirected(z['ei'], None, N), O.is_undirected(z['und_ei'], None, N), O.is_undirected(z['und_ei'], z['und_attr'], N)]
    assert got == z['is_und'].tolist()
