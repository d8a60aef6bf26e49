"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (its own CPU operators compiled unmodified into
oracle/_ref/_torch_ext.so by oracle/build_ref.py).  Run in the authoring container only:

    python tests/golden/make_golden.py

/root/reference does not exist on the GPU box, so the vectors are committed.  Contents:

  kat_*.npz      the reference's own known-answer tests, inputs AND the literal expectations written in
                 /root/reference/tests/mpops/torch_ops.py:18-72, tests/layers/conv/test_message_passing.py:12-24,
                 tests/utils/test_softmax.py:13-20, tests/utils/test_degree.py:5-9 and the docstring example of
                 gammagl/mpops/torch.py:322-330; the script asserts the compiled reference reproduces every one.
  rand_*.npz     seeded randomized cases (forward + backward of all 7 c_* operators) produced by the compiled
                 reference: unsorted ids, empty segments, duplicate edges, ties (integer-valued floats),
                 K in {1, 3, 16, 47, 100}.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
from oracle.build_ref import load_ref  # noqa: E402

ref = load_ref()
assert ref is not None, 'reference extension could not be built (needs /root/reference)'


def t(a, dtype=None):
    return torch.tensor(a, dtype=dtype)


# --------------------------------------------------------------------------------------------------
# 1. known-answer tests of the reference
# --------------------------------------------------------------------------------------------------
def kat_torch_ops():
    """tests/mpops/torch_ops.py: 3 ops x 7 dtypes x rank 1/2/3, index [0,0,1], N=2."""
    dtypes = [torch.int8, torch.int16, torch.int32, torch.int64, torch.float16, torch.float32, torch.float64]
    gen = {1: [1, 2, 3], 2: [[1, 2], [3, 4], [5, 6]], 3: [[[1, 2], [3, 4]], [[5, 6], [7, 8]], [[9, 10], [11, 12]]]}
    exp = {
        'max': {1: [2, 3], 2: [[3, 4], [5, 6]], 3: [[[5, 6], [7, 8]], [[9, 10], [11, 12]]]},
        'sum': {1: [3, 3], 2: [[4, 6], [5, 6]], 3: [[[6, 8], [10, 12]], [[9, 10], [11, 12]]]},
        'mean': {1: [1.5, 3], 2: [[2, 3], [5, 6]], 3: [[[3, 4], [5, 6]], [[9, 10], [11, 12]]]},
    }
    fn = {'max': ref.c_segment_max, 'sum': ref.c_segment_sum, 'mean': ref.c_segment_mean}
    index = torch.tensor([0, 0, 1], dtype=torch.int64)
    out = {}
    for op in ('max', 'sum', 'mean'):
        for dt in dtypes:
            for dim in (1, 2, 3):
                x = torch.tensor(gen[dim], dtype=dt)
                # the reference builds its expectation with convert_to_tensor(..., dtype) => 1.5 truncates for ints
                e = torch.tensor(exp[op][dim], dtype=torch.float64).to(dt)
                got = fn[op](x, index, 2)
                assert np.allclose(got.double().numpy(), e.double().numpy(), atol=1e-5), (op, dt, dim, got, e)
                key = '%s_%s_%d' % (op, str(dt).split('.')[-1], dim)
                out[key + '_x'] = x.numpy()
                out[key + '_expected'] = e.numpy()
    out['index'] = index.numpy()
    np.savez(os.path.join(HERE, 'kat_torch_ops.npz'), **out)
    print('kat_torch_ops: %d arrays' % len(out))


def kat_message_passing():
    """tests/layers/conv/test_message_passing.py (bare MessagePassing -> gspmm with unit weights),
    tests/utils/test_softmax.py, tests/utils/test_degree.py, gspmm docstring (mpops/torch.py:322-330)."""
    out = {}
    x = torch.arange(0, 8, dtype=torch.float32).reshape(4, 2)
    ei = torch.tensor([[0, 1, 2, 2, 3], [1, 2, 1, 0, 3]])
    w = torch.ones(5)
    exp = {'sum': [[4., 5.], [4., 6.], [2., 3.], [6., 7.]], 'mean': [[4., 5.], [2., 3.], [2., 3.], [6., 7.]],
           'max': [[4., 5.], [4., 5.], [2., 3.], [6., 7.]]}
    for op, f in (('sum', ref.c_spmm_sum), ('mean', ref.c_spmm_mean), ('max', ref.c_spmm_max)):
        got = f(ei, w, x)
        assert got.tolist() == exp[op], (op, got)
        out['mp_' + op] = np.array(exp[op], dtype=np.float32)
    out['mp_x'] = x.numpy()
    out['mp_edge_index'] = ei.numpy()

    # softmax KAT: row 2 of the score ~ [0.7311, 0.9526]
    xs = torch.tensor([[1, 1], [1, 1], [2, 4], [2, 4]], dtype=torch.float32)
    xe = xs[ei[0]]
    mx = ref.c_segment_max(xe, ei[1], 4)
    ex = torch.exp(xe - mx[ei[1]])
    den = ref.c_segment_sum(ex, ei[1], 4)
    score = ex / (den[ei[1]] + 1e-16)
    assert abs(float((score[2] - torch.tensor([0.7311, 0.9526])).sum())) < 1e-4
    out['softmax_data'] = xe.numpy()
    out['softmax_ids'] = ei[1].numpy()
    out['softmax_score'] = score.numpy()
    out['softmax_row2_expected'] = np.array([0.7311, 0.9526], dtype=np.float32)

    # degree KAT
    row = torch.tensor([0, 1, 0, 2, 0])
    deg = ref.c_segment_sum(torch.ones(5, dtype=torch.int64), row, 3)
    assert deg.tolist() == [3, 1, 1] and deg.dtype == torch.int64
    out['degree_index'] = row.numpy()
    out['degree_expected'] = np.array([3, 1, 1], dtype=np.int64)

    # gspmm docstring
    idx = torch.tensor([[0, 1, 1, 1, 2, 3, 3, 4], [1, 0, 2, 3, 1, 1, 4, 3]])
    wd = 2 * torch.ones(8)
    xd = 2 * torch.ones(5, 8)
    got = ref.c_spmm_sum(idx, wd, xd)
    e = np.array([[4.] * 8, [12.] * 8, [4.] * 8, [8.] * 8, [4.] * 8], dtype=np.float32)
    assert np.array_equal(got.numpy(), e)
    out['doc_index'] = idx.numpy()
    out['doc_weight'] = wd.numpy()
    out['doc_x'] = xd.numpy()
    out['doc_expected'] = e

    # unsorted_segment_* docstring (mpops/torch.py:64-71,120-127,180-187)
    xs3 = torch.tensor([[1., 2., 3., 4.], [4., 3., 2., 1.], [5., 6., 7., 8.]])
    ids3 = torch.tensor([0, 2, 0])
    out['doc_seg_x'] = xs3.numpy()
    out['doc_seg_ids'] = ids3.numpy()
    out['doc_seg_sum'] = ref.c_segment_sum(xs3, ids3, 3).numpy()
    assert out['doc_seg_sum'].tolist() == [[6., 8., 10., 12.], [0., 0., 0., 0.], [4., 3., 2., 1.]]
    out['doc_seg_mean'] = ref.c_segment_mean(xs3, ids3, 3).numpy()
    assert out['doc_seg_mean'].tolist() == [[3., 4., 5., 6.], [0., 0., 0., 0.], [4., 3., 2., 1.]]
    out['doc_seg_max_ext'] = ref.c_segment_max(xs3, ids3, 3).numpy()   # empty row = lowest() in the ext
    np.savez(os.path.join(HERE, 'kat_message_passing.npz'), **out)
    print('kat_message_passing: %d arrays' % len(out))


# --------------------------------------------------------------------------------------------------
# 2. randomized vectors from the compiled reference
# --------------------------------------------------------------------------------------------------
def rand_case(name, N, E, K, seed, ties=False, H=None):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, N, (E,), generator=g)
    # leave ~10% of the nodes without in-edges (empty segments) and create duplicates
    dst = torch.randint(0, max(1, int(N * 0.9)), (E,), generator=g)
    dst[: E // 8] = dst[E // 8: 2 * (E // 8)]
    src[: E // 16] = src[E // 8: E // 8 + E // 16]
    index = torch.stack([src, dst])
    if ties:
        x = torch.randint(-3, 4, (N, K), generator=g).float()
        w = torch.randint(1, 3, (E,), generator=g).float()
        msg = torch.randint(-3, 4, (E, K), generator=g).float()
    else:
        x = torch.randn(N, K, generator=g)
        w = torch.rand(E, generator=g) + 0.1
        msg = torch.randn(E, K, generator=g)
    gout = torch.randn(N, K, generator=g)
    out = {'index': index.numpy(), 'x': x.numpy(), 'w': w.numpy(), 'msg': msg.numpy(), 'gout': gout.numpy(),
           'N': np.int64(N)}

    for op, f in (('sum', ref.c_segment_sum), ('mean', ref.c_segment_mean), ('max', ref.c_segment_max)):
        m = msg.clone().requires_grad_(True)
        o = f(m, dst, N)
        out['seg_%s_out' % op] = o.detach().numpy()
        if op == 'max':
            # gradient of empty segments leaks into row N in the reference (src/segment_max.cpp:55-59): mask them
            nonempty = torch.zeros(N, dtype=torch.bool)
            nonempty[dst] = True
            go = gout * nonempty[:, None]
        else:
            go = gout
        o.backward(go)
        out['seg_%s_gx' % op] = m.grad.numpy()
    for op, f in (('sum', ref.c_spmm_sum), ('mean', ref.c_spmm_mean), ('max', ref.c_spmm_max)):
        xx = x.clone().requires_grad_(True)
        o = f(index, w, xx)
        out['spmm_%s_out' % op] = o.detach().numpy()
        o.backward(gout)
        out['spmm_%s_gx' % op] = xx.grad.numpy()
    if H is not None and K % H == 0:
        C = K // H
        wh = (torch.rand(E, H, generator=g) + 0.1).requires_grad_(True)
        xh = x.clone().view(N, H, C).requires_grad_(True)
        o = ref.c_bspmm_sum(index, wh, xh)
        out['bspmm_w'] = wh.detach().numpy()
        out['bspmm_out'] = o.detach().numpy()
        o.backward(gout.view(N, H, C))
        out['bspmm_gx'] = xh.grad.numpy()
        out['bspmm_gw'] = wh.grad.numpy()
        out['H'] = np.int64(H)
    np.savez_compressed(os.path.join(HERE, 'rand_%s.npz' % name), **out)
    print('rand_%s: N=%d E=%d K=%d' % (name, N, E, K))


if __name__ == '__main__':
    kat_torch_ops()
    kat_message_passing()
    rand_case('k1', N=200, E=1500, K=1, seed=1)
    rand_case('k3', N=150, E=1200, K=3, seed=2)
    rand_case('k16_h4', N=300, E=3000, K=16, seed=3, H=4)
    rand_case('k47', N=120, E=900, K=47, seed=4)
    rand_case('k100', N=150, E=1000, K=100, seed=5)
    rand_case('k64_ties_h8', N=120, E=1500, K=64, seed=6, ties=True, H=8)
    rand_case('k128', N=24, E=1500, K=128, seed=7)    # average degree ~70: rows longer than one 32-edge batch
