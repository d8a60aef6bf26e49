"""GPU tests of the sort-free segment reduce (csrc/oneshot.cu, gglb200_segment_reduce) and of the policy that picks it:
  * max (+ int64 arg) bit-exact against the oracle, ties, -0.0 / +0.0, NaN, empty rows, hubs;
  * sum / mean within 1e-6 of the oracle, exact on integer-valued data (the reference KATs, degree());
  * backward through autograd without a CSR (row gather, counts from the pass itself);
  * ids outside [0, N) raise; the policy builds the CSR from the second sighting on."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def mp():
    from gammagl_b200 import _lib, mpops
    _lib.load()
    return mpops


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


@pytest.mark.segment_mode('oneshot')
def test_reference_kats_through_the_oneshot_path(mp):
    """/root/reference/tests/mpops/torch_ops.py:18-72 (float32 cases; the other dtypes take the CSR path by design)."""
    z = np.load(os.path.join(GOLDEN, 'kat_torch_ops.npz'))
    idx = torch.from_numpy(z['index']).to(DEV)
    fn = {'max': mp.segment_max, 'sum': mp.segment_sum, 'mean': mp.segment_mean}
    n = 0
    for key in z.files:
        if not key.endswith('_x') or z[key].dtype != np.float32:
            continue
        exp = z[key[:-2] + '_expected']
        got = fn[key.split('_')[0]](torch.from_numpy(z[key]).to(DEV), idx, 2).cpu().numpy()
        assert got.shape == exp.shape and np.allclose(got, exp, atol=1e-6), key
        n += 1
    assert n >= 9


@pytest.mark.segment_mode('oneshot')
@pytest.mark.parametrize('E,N,K,hub', [(1 << 16, 1 << 12, 100, 0), (50000, 3000, 33, 4000), (20000, 500, 4, 3000), (30000, 2500, 1, 0),
                                        (9000, 300, 130, 0), (40000, 700, 2, 9000), (25000, 1200, 6, 0)])
def test_oneshot_vs_oracle(mp, E, N, K, hub):
    g = torch.Generator().manual_seed(E + K)
    ids = torch.randint(0, max(1, N - N // 16), (E,), generator=g)              # last 1/16 of the segments stay empty
    if hub:
        ids[:hub] = 3
    msg = torch.randn(E, K, generator=g)
    msg[: E // 4] = torch.randint(-2, 3, (E // 4, K), generator=g).float()      # ties for max
    msg[E // 4: E // 4 + 50] = 0.0
    msg[E // 4: E // 4 + 25] *= -1.0                                            # -0.0 next to +0.0
    gout = torch.randn(N, K, generator=g)
    nonempty = torch.from_numpy(np.bincount(ids.numpy(), minlength=N) > 0)
    fn = {'sum': mp.unsorted_segment_sum, 'mean': mp.unsorted_segment_mean, 'max': mp.unsorted_segment_max}
    for op in ('sum', 'mean', 'max'):
        r = O.segment(op, msg.numpy(), ids.numpy(), N)
        ref, arg = (r if op == 'max' else (r, None))
        m = msg.to(DEV).requires_grad_(True)
        out = fn[op](m, ids.to(DEV), N)
        got = out.detach().cpu().numpy()
        if op == 'max':
            assert np.array_equal(got, ref)                                      # bit-exact, incl. the sign of zero and lowest()
        else:
            assert rel(got, ref) < 2e-6
        go = gout * nonempty[:, None] if op == 'max' else gout
        out.backward(go.to(DEV))
        gref = O.segment_bwd(op, go.numpy(), ids.numpy(), E, arg)
        assert np.array_equal(m.grad.cpu().numpy(), gref), (op, 'bwd')


@pytest.mark.segment_mode('oneshot')
def test_oneshot_max_arg_and_nan(mp):
    N, E, K = 700, 30000, 12
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, N - 40, (E,), generator=g)
    ids[:5000] = 9
    msg = torch.randint(-2, 3, (E, K), generator=g).float()
    ref, arg_ref = O.segment('max', msg.numpy(), ids.numpy(), N)
    out, arg, cnt = mp.segment_reduce_oneshot(msg.to(DEV), ids.to(DEV), N, 2)
    assert arg.dtype == torch.int64 and np.array_equal(arg.cpu().numpy(), arg_ref)      # first maximal edge, sentinel N for empty rows
    assert np.array_equal(out.cpu().numpy(), ref)
    assert np.array_equal(cnt.cpu().numpy(), np.bincount(ids.numpy(), minlength=N))
    # NaN never replaces the running maximum (`out < x` is false for NaN): the result is the maximum of the other messages
    m2 = msg.clone()
    m2[::7] = float('nan')
    ref2, arg2 = O.segment('max', m2.numpy(), ids.numpy(), N)
    out2, a2, _ = mp.segment_reduce_oneshot(m2.to(DEV), ids.to(DEV), N, 2)
    assert np.array_equal(out2.cpu().numpy(), ref2) and np.array_equal(a2.cpu().numpy(), arg2)


@pytest.mark.segment_mode('fast')
def test_policy_degree_and_second_sighting(mp):
    from gammagl_b200 import csr
    csr.clear_caches()
    N, E = 5000, 200000
    g = torch.Generator().manual_seed(8)
    ids = torch.randint(0, N, (E,), generator=g).to(DEV)
    launches = []
    real = mp.segment_reduce_oneshot
    mp.segment_reduce_oneshot = lambda *a: (launches.append(a[3]), real(*a))[1]
    try:
        deg = mp.unsorted_segment_sum(torch.ones(E, device=DEV), ids, N)          # K = 1: always one-shot, exact counts
        assert launches == [0]
        assert np.array_equal(deg.cpu().numpy(), np.bincount(ids.cpu().numpy(), minlength=N).astype(np.float32))
        x = torch.randn(E, 32, generator=g).to(DEV)
        a = mp.unsorted_segment_max(x, ids, N)                                     # wide max: never one-shot (64-bit atomics per element)
        assert launches == [0] and csr.segment_csr_if_cached(ids, N) is not None
        csr.clear_caches()
        s1 = mp.unsorted_segment_sum(x, ids, N)                                    # fast mode, first sighting of these ids: one-shot
        assert launches == [0, 0] and csr.segment_csr_if_cached(ids, N) is None
        s2 = mp.unsorted_segment_sum(x, ids, N)                                    # second sighting: the CSR is built and used
        assert launches == [0, 0] and csr.segment_csr_if_cached(ids, N) is not None
        assert rel(s1.cpu().numpy(), s2.cpu().numpy()) < 2e-6
        old = mp._SEGMENT_MODE
        mp._SEGMENT_MODE = 'auto'                                                  # auto: wide sums never leave the bit-exact path
        try:
            csr.clear_caches()
            mp.unsorted_segment_sum(x, ids, N)
            assert launches == [0, 0]
            mp.unsorted_segment_max(torch.randn(E, 2, generator=g).to(DEV), ids, N)    # 8-byte rows: one-shot (max is exact there)
            assert launches == [0, 0, 2]
        finally:
            mp._SEGMENT_MODE = old
        with torch.no_grad():
            bad = ids.clone()
            bad[5] = N
        with pytest.raises(IndexError):
            mp.unsorted_segment_sum(torch.ones(E, device=DEV), bad, N)
        assert a.shape == (N, 32)
    finally:
        mp.segment_reduce_oneshot = real
        csr.clear_caches()
