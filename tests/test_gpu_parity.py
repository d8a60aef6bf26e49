"""GPU parity tests (-m gpu): the CUDA path, called through gammagl_b200.mpops -> C ABI (libgglb200.so), against
  * the reference's own known-answer vectors (tests/golden/kat_*.npz),
  * vectors produced by the compiled reference (tests/golden/rand_*.npz),
  * the CPU oracle (oracle/) on seeded inputs at sizes it finishes in seconds,
  * size-independent properties at BASELINE.json's full sizes (test_gpu_fullsize.py).

Bars: integer / index results bit-exact; fp32 sums BIT-EXACT for rows that are not split (<= 1024 edges; same
serial order and rounding as the reference loop) and within 1e-5 relative otherwise; bf16 within 1e-2 of the
fp32 oracle on up-cast inputs (north_star tolerances).
"""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RAND = sorted(glob.glob(os.path.join(GOLDEN, 'rand_*.npz')))
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def mp():
    from gammagl_b200 import _lib, mpops
    _lib.load()  # must raise if the native library is missing: no fallback
    return mpops


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


# ------------------------------------------------------------------------------------------------------
# reference KATs
# ------------------------------------------------------------------------------------------------------
def test_kat_torch_ops(mp):
    """/root/reference/tests/mpops/torch_ops.py:18-72, 63 cases, every dtype of the reference's dispatch."""
    z = np.load(os.path.join(GOLDEN, 'kat_torch_ops.npz'))
    idx = cu(z['index'])
    fn = {'max': mp.segment_max, 'sum': mp.segment_sum, 'mean': mp.segment_mean}
    n = 0
    for key in z.files:
        if not key.endswith('_x'):
            continue
        op = key.split('_')[0]
        exp = z[key[:-2] + '_expected']
        got = fn[op](cu(z[key]), idx, 2).cpu().numpy()
        assert got.dtype == exp.dtype and got.shape == exp.shape, key
        if np.issubdtype(exp.dtype, np.integer):
            assert np.array_equal(got, exp), key
        else:
            assert np.allclose(got.astype(np.float64), exp.astype(np.float64), atol=1e-5), key
        n += 1
    assert n == 63


def test_kat_message_passing_softmax_degree_docstrings(mp):
    z = np.load(os.path.join(GOLDEN, 'kat_message_passing.npz'))
    ei, x = cu(z['mp_edge_index']), cu(z['mp_x'])
    for op in ('sum', 'mean', 'max'):
        assert mp.gspmm(ei, None, x, op).cpu().tolist() == z['mp_' + op].tolist(), op
    score = mp.segment_softmax(cu(z['softmax_data']), cu(z['softmax_ids']), 4).cpu().numpy()
    assert abs(float((score[2] - z['softmax_row2_expected']).sum())) < 1e-4
    assert np.allclose(score, z['softmax_score'], rtol=1e-5, atol=1e-7)
    deg = mp.unsorted_segment_sum(torch.ones(5, dtype=torch.int64, device=DEV), cu(z['degree_index']), 3)
    assert deg.dtype == torch.int64 and deg.cpu().tolist() == z['degree_expected'].tolist()
    out = mp.gspmm(cu(z['doc_index']), cu(z['doc_weight']), cu(z['doc_x']))
    assert np.array_equal(out.cpu().numpy(), z['doc_expected'])
    sx, sid = cu(z['doc_seg_x']), cu(z['doc_seg_ids'])
    assert np.array_equal(mp.unsorted_segment_sum(sx, sid, 3).cpu().numpy(), z['doc_seg_sum'])
    assert np.array_equal(mp.unsorted_segment_mean(sx, sid, 3).cpu().numpy(), z['doc_seg_mean'])
    assert np.array_equal(mp.unsorted_segment_max(sx, sid, 3).cpu().numpy(), z['doc_seg_max_ext'])
    # num_segments=None -> max(ids)+1 (mpops/torch.py:74-75)
    assert mp.unsorted_segment_sum(sx, sid).shape[0] == 3


# ------------------------------------------------------------------------------------------------------
# vectors produced by the compiled reference: forward AND backward, bit-exact
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('path', RAND, ids=[os.path.basename(p)[:-4] for p in RAND])
def test_reference_vectors_bit_exact(mp, path):
    z = np.load(path)
    N = int(z['N'])
    index, x, w, msg, gout = cu(z['index']), cu(z['x']), cu(z['w']), cu(z['msg']), cu(z['gout'])
    dst = index[1]
    E = index.shape[1]
    nonempty = torch.zeros(N, dtype=torch.bool, device=DEV)
    nonempty[dst] = True
    seg = {'sum': mp.unsorted_segment_sum, 'mean': mp.unsorted_segment_mean, 'max': mp.unsorted_segment_max}
    for op in ('sum', 'mean', 'max'):
        m = msg.clone().requires_grad_(True)
        o = seg[op](m, dst, N)
        assert np.array_equal(o.detach().cpu().numpy(), z['seg_%s_out' % op]), ('segment fwd', op)
        o.backward(gout * nonempty[:, None] if op == 'max' else gout)
        got, ref = m.grad.cpu().numpy(), z['seg_%s_gx' % op]
        if op == 'max' and not bool(nonempty.all()) and N < E:
            keep = np.ones(E, dtype=bool)   # row N of the reference is corrupted by its sentinel collision
            keep[N] = False                 # (src/segment_max.cpp:55-59), see tests/test_oracle.py
            got, ref = got[keep], ref[keep]
        assert np.array_equal(got, ref), ('segment bwd', op)

        xx = x.clone().requires_grad_(True)
        o = mp.gspmm(index, w, xx, op)
        assert np.array_equal(o.detach().cpu().numpy(), z['spmm_%s_out' % op]), ('spmm fwd', op)
        o.backward(gout)
        assert np.array_equal(xx.grad.cpu().numpy(), z['spmm_%s_gx' % op]), ('spmm bwd', op)
    if 'bspmm_out' in z.files:
        H = int(z['H'])
        C = x.shape[1] // H
        wh = cu(z['bspmm_w']).requires_grad_(True)
        xh = x.clone().view(N, H, C).requires_grad_(True)
        o = mp.bspmm(index, wh, xh)
        assert np.array_equal(o.detach().cpu().numpy(), z['bspmm_out'])
        o.backward(gout.view(N, H, C))
        assert np.array_equal(xh.grad.cpu().numpy(), z['bspmm_gx'])
        assert rel_err(wh.grad.cpu().numpy(), z['bspmm_gw']) < 1e-5   # cross-lane tree sum over C


# ------------------------------------------------------------------------------------------------------
# oracle comparison at sizes the oracle finishes in seconds; power-law rows exercise the split path
# ------------------------------------------------------------------------------------------------------
def make_graph(N, E, seed, hub=0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, N, (E,), generator=g)
    dst = torch.randint(0, max(1, N - N // 16), (E,), generator=g)      # last 1/16 of the nodes: no in-edges
    if hub:
        dst[:hub] = 3                                                   # one row far above GGLB200_LONG_ROW
        src[hub:2 * hub] = 5                                            # and one long row in the transposed CSR
    perm = torch.randperm(E, generator=g)
    return torch.stack([src[perm], dst[perm]])


@pytest.mark.parametrize('N,E,K,hub', [(1 << 12, 1 << 16, 100, 0), (1 << 12, 1 << 16, 128, 5000), (3000, 40000, 47, 3000),
                                        (1 << 10, 1 << 14, 602, 0), (777, 9000, 7, 2500), (5000, 60000, 1, 4000),
                                        (2000, 30000, 16, 0), (1500, 20000, 64, 1500), (600, 5000, 2, 0)])
def test_gspmm_vs_oracle(mp, N, E, K, hub):
    index = make_graph(N, E, seed=N + K, hub=hub)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(N, K, generator=g)
    w = torch.rand(E, generator=g) + 0.05
    gout = torch.randn(N, K, generator=g)
    deg_in = np.bincount(index[1].numpy(), minlength=N)
    deg_out = np.bincount(index[0].numpy(), minlength=N)
    short_in, short_out = deg_in <= 1024, deg_out <= 1024
    for op in ('sum', 'mean', 'max'):
        ref, aux = O.spmm_fwd(op, index.numpy(), w.numpy(), x.numpy())
        xx = x.to(DEV).requires_grad_(True)
        out = mp.gspmm(index.to(DEV), w.to(DEV), xx, op)
        got = out.detach().cpu().numpy()
        assert np.array_equal(got[short_in], ref[short_in]), (op, 'unsplit rows must be bit-exact')
        if op == 'max':
            assert np.array_equal(got, ref), 'max is exact for every row'
        else:
            assert rel_err(got, ref) < 1e-5, op
        out.backward(gout.to(DEV))
        gref = O.spmm_bwd(op, index.numpy(), w.numpy(), gout.numpy(), aux)
        ggot = xx.grad.cpu().numpy()
        if op == 'max':
            assert np.array_equal(ggot, gref)
        else:
            assert np.array_equal(ggot[short_out], gref[short_out]), (op, 'bwd unsplit rows bit-exact')
            assert rel_err(ggot, gref) < 1e-5, op


def test_odd_feature_widths_padded_and_unpadded_agree_bitwise(mp, monkeypatch):
    """fp32 rows that are not a multiple of 16 bytes (47 classes) are padded to 48 columns on the host side of the aggregation
    (16-byte lanes instead of one element per lane); columns are independent, so both forms must agree bit for bit -- values,
    arg-max indices and gradients."""
    N, E, K = 2500, 40000, 47
    index = make_graph(N, E, seed=3, hub=2000).to(DEV)
    g = torch.Generator().manual_seed(23)
    x = torch.randn(N, K, generator=g).to(DEV)
    w = (torch.rand(E, generator=g) + 0.05).to(DEV)
    gout = torch.randn(N, K, generator=g).to(DEV)
    res = {}
    for pad in (True, False):
        monkeypatch.setattr(mp, '_PAD_ODD_K', pad)
        for op in ('sum', 'mean', 'max'):
            xx = x.clone().requires_grad_(True)
            out = mp.gspmm(index, w, xx, op)
            out.backward(gout)
            res[(pad, op)] = (out.detach(), xx.grad)
    for op in ('sum', 'mean', 'max'):
        assert torch.equal(res[(True, op)][0], res[(False, op)][0]), op
        assert torch.equal(res[(True, op)][1], res[(False, op)][1]), op


@pytest.mark.parametrize('E,N,K,hub', [(1 << 16, 1 << 12, 100, 0), (50000, 3000, 33, 4000), (20000, 500, 4, 3000),
                                        (30000, 2500, 1, 0), (9000, 300, 130, 0)])
def test_segment_ops_vs_oracle(mp, E, N, K, hub):
    index = make_graph(N, E, seed=E + K, hub=hub)
    ids = index[1]
    g = torch.Generator().manual_seed(11)
    msg = torch.randn(E, K, generator=g)
    msg[: E // 4] = torch.randint(-2, 3, (E // 4, K), generator=g).float()   # ties for max
    gout = torch.randn(N, K, generator=g)
    deg = np.bincount(ids.numpy(), minlength=N)
    short = deg <= 1024
    nonempty = torch.from_numpy(deg > 0)
    fn = {'sum': mp.unsorted_segment_sum, 'mean': mp.unsorted_segment_mean, 'max': mp.unsorted_segment_max}
    for op in ('sum', 'mean', 'max'):
        r = O.segment(op, msg.numpy(), ids.numpy(), N)
        ref, arg = (r if op == 'max' else (r, None))
        m = msg.to(DEV).requires_grad_(True)
        out = fn[op](m, ids.to(DEV), N)
        got = out.detach().cpu().numpy()
        assert np.array_equal(got[short], ref[short]), op
        if op == 'max':
            assert np.array_equal(got, ref)
        else:
            assert rel_err(got, ref) < 1e-5
        go = gout * nonempty[:, None] if op == 'max' else gout
        out.backward(go.to(DEV))
        gref = O.segment_bwd(op, go.numpy(), ids.numpy(), E, arg)
        assert np.array_equal(m.grad.cpu().numpy(), gref), (op, 'bwd')


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16, torch.float64, torch.int8])
@pytest.mark.parametrize('E,N,K,hub', [(70000, 3000, 32, 5000), (5000, 7000, 3, 0), (33333, 100, 100, 0), (1000, 10, 1, 900),
                                        (40000, 2000, 260, 0)])
def test_segment_backward_scatter_and_gather_forms_agree(mp, dtype, E, N, K, hub):
    """gglb200_segment_expand_bwd (scatter through the CSR) == gglb200_segment_gather_bwd (row gather), bit for bit."""
    from gammagl_b200 import _lib
    from gammagl_b200.csr import _ptr, _stream, segment_csr_of
    lib = _lib.load()
    ids = make_graph(N, E, seed=E + K, hub=hub)[1].to(DEV)
    csr = segment_csr_of(ids, N)
    g = (torch.randn(N, K, generator=torch.Generator().manual_seed(2)) * 8).to(dtype).to(DEV)
    code = mp._dtype_code(g)
    for reduce in (_lib.GGL_SUM, _lib.GGL_MEAN):
        if reduce == _lib.GGL_MEAN and not dtype.is_floating_point:
            continue
        a = torch.full((E, K), 77, dtype=dtype, device=DEV)
        b = torch.full((E, K), 55, dtype=dtype, device=DEV)
        _lib.check(lib.gglb200_segment_gather_bwd(reduce, _ptr(g), _ptr(ids), _ptr(csr.rowptr), _ptr(a), E, K, code,
                                                  _stream(g.device)), 'gather')
        _lib.check(lib.gglb200_segment_expand_bwd(reduce, _ptr(g), _ptr(csr.rowptr), _ptr(csr.perm), _ptr(csr.plan), _ptr(b),
                                                  E, N, K, code, _stream(g.device)), 'expand')
        assert torch.equal(a, b), reduce
        if reduce == _lib.GGL_SUM:
            assert torch.equal(a, g[ids])


def test_segment_max_arg_indices_bit_exact(mp):
    """arg-max edge ids (first edge wins on ties; sentinel N for empty segments) must equal the oracle's exactly."""
    from gammagl_b200 import _lib
    from gammagl_b200.csr import segment_csr_of
    N, E, K = 900, 30000, 12
    index = make_graph(N, E, seed=5, hub=2000)
    ids = index[1].to(DEV)
    msg = torch.randint(-2, 3, (E, K), generator=torch.Generator().manual_seed(3)).float()
    _, arg_ref = O.segment('max', msg.numpy(), index[1].numpy(), N)
    csr = segment_csr_of(ids, N)
    out, arg = mp.spmm_csr(csr, csr.perm, msg.to(DEV), _lib.GGL_MAX, want_arg=True, arg_empty=N)
    assert arg.dtype == torch.int64
    assert np.array_equal(arg.cpu().numpy(), arg_ref)


@pytest.mark.parametrize('N,E,H,C,hub', [(2000, 40000, 8, 8, 0), (1500, 30000, 8, 64, 1500), (900, 9000, 4, 16, 0),
                                          (500, 6000, 3, 5, 0), (700, 8000, 1, 32, 1200), (400, 5000, 2, 128, 0),
                                          # head widths that are not a multiple of the 16-byte lane: lanes span two heads
                                          (1200, 30000, 4, 25, 0), (1200, 30000, 4, 25, 2500), (800, 12000, 3, 12, 0),
                                          (800, 12000, 6, 7, 1500), (600, 9000, 5, 6, 0), (600, 9000, 2, 50, 0)])
def test_bspmm_vs_oracle(mp, N, E, H, C, hub):
    index = make_graph(N, E, seed=H * 100 + C, hub=hub)
    g = torch.Generator().manual_seed(13)
    x = torch.randn(N, H, C, generator=g)
    w = torch.rand(E, H, generator=g)
    gout = torch.randn(N, H, C, generator=g)
    ref = O.bspmm_fwd(index.numpy(), w.numpy(), x.numpy())
    gx_ref, gw_ref = O.bspmm_bwd(index.numpy(), w.numpy(), x.numpy(), gout.numpy())
    xx = x.to(DEV).requires_grad_(True)
    ww = w.to(DEV).requires_grad_(True)
    out = mp.bspmm(index.to(DEV), ww, xx)
    assert rel_err(out.detach().cpu().numpy(), ref) < 1e-5
    if hub == 0:
        assert np.array_equal(out.detach().cpu().numpy(), ref)
    out.backward(gout.to(DEV))
    assert rel_err(xx.grad.cpu().numpy(), gx_ref) < 1e-5
    assert rel_err(ww.grad.cpu().numpy(), gw_ref) < 1e-5


@pytest.mark.parametrize('path', ['csr', 'oneshot'])
@pytest.mark.parametrize('H', [1, 4, 8, 13])
def test_segment_softmax_vs_oracle(mp, H, path, monkeypatch):
    """Both forms of the edge softmax: statistics over the segment CSR (serial-order sums) and the sort-free streaming passes."""
    monkeypatch.setattr(mp, '_SOFTMAX_PATH', path)
    N, E = 1200, 30000
    index = make_graph(N, E, seed=H, hub=2000)
    g = torch.Generator().manual_seed(17)
    logit = torch.randn(E, H, generator=g) * 3
    logit[:200] = 9000.0 + torch.randn(200, H, generator=g)          # huge, negative and constant logits: the exact maximum matters
    logit[200:400] = -7000.0
    ref = O.segment_softmax(logit.numpy(), index[1].numpy(), N)
    ll = logit.to(DEV).requires_grad_(True)
    got = mp.segment_softmax(ll, index[1].to(DEV), N)
    assert rel_err(got.detach().cpu().numpy(), ref) < 1e-5
    # rows sum to one
    s = torch.zeros(N, H, device=DEV).index_add_(0, index[1].to(DEV), got.detach())
    assert torch.allclose(s[s > 0], torch.ones_like(s[s > 0]), atol=1e-5)
    # backward against autograd of the composed definition (torch ops, fp64)
    ga = torch.randn(E, H, generator=g)
    got.backward(ga.to(DEV))
    l64 = logit.double().requires_grad_(True)
    mx = torch.full((N, H), -1e300, dtype=torch.float64).scatter_reduce(0, index[1][:, None].expand(E, H), l64, 'amax')
    ex = torch.exp(l64 - mx[index[1]].detach())
    den = torch.zeros(N, H, dtype=torch.float64).index_add_(0, index[1], ex)
    (ex / (den[index[1]] + 1e-16)).backward(ga.double())
    assert rel_err(ll.grad.cpu().numpy(), l64.grad.numpy()) < 1e-4


@pytest.mark.parametrize('dtype,tol', [(torch.bfloat16, 1e-2), (torch.float16, 2e-3)])
@pytest.mark.parametrize('K', [100, 128, 64, 6, 47])
def test_low_precision_vs_fp32_oracle(mp, dtype, tol, K):
    """bf16/f16 storage, fp32 accumulate, compared with the fp32 oracle on the up-cast inputs (north_star: 1e-2)."""
    N, E = 3000, 60000
    index = make_graph(N, E, seed=K, hub=1500)
    g = torch.Generator().manual_seed(19)
    x = torch.randn(N, K, generator=g).to(dtype)
    w = torch.rand(E, generator=g)
    for op in ('sum', 'mean', 'max'):
        ref, _ = O.spmm_fwd(op, index.numpy(), w.numpy(), x.float().numpy())
        got = mp.gspmm(index.to(DEV), w.to(DEV), x.to(DEV), op)
        assert got.dtype == dtype
        got = got.float().cpu().numpy()
        mask = ref > -1e30                                      # empty rows of max: lowest() of the storage type
        assert rel_err(got[mask], ref[mask]) < tol, (op, dtype)
    msg = torch.randn(E, K, generator=g).to(dtype)
    ref = O.segment('sum', msg.float().numpy(), index[1].numpy(), N)
    got = mp.unsorted_segment_sum(msg.to(DEV), index[1].to(DEV), N).float().cpu().numpy()
    assert rel_err(got, ref) < tol


def test_csr_build_matches_stable_argsort_and_ind2ptr(mp):
    """COO->CSR ingest vs numpy stable argsort + the oracle's ind2ptr (storage.py:124-140, convert.cpp:58-90)."""
    from gammagl_b200.csr import build_csr
    N, E = 5000, 200000
    index = make_graph(N, E, seed=23, hub=3000)
    c = build_csr(index[1].to(DEV), index[0].to(DEV), N)
    perm_ref = np.argsort(index[1].numpy(), kind='stable')
    assert np.array_equal(c.perm.cpu().numpy(), perm_ref.astype(np.int32))
    assert np.array_equal(c.col.cpu().numpy(), index[0].numpy()[perm_ref].astype(np.int32))
    assert np.array_equal(c.rowptr.cpu().numpy().astype(np.int64), O.ind2ptr(np.sort(index[1].numpy()), N))
    assert c.has_long
    plan = c.plan.cpu().numpy()
    deg = np.bincount(index[1].numpy(), minlength=N)
    assert plan[1] == int((deg > 1024).sum())
    assert plan[0] == int(np.ceil(deg[deg > 1024] / 512).sum())


def test_edge_cases(mp):
    # empty x ("rgcn empty segment_ids", mpops/torch.py:76-78; cpu/segment_sum_cpu.cpp:28-30)
    out = mp.unsorted_segment_sum(torch.zeros(0, 4, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV), 3)
    assert out.shape == (3, 4) and float(out.abs().sum()) == 0
    # graph without edges
    ei = torch.zeros(2, 0, dtype=torch.int64, device=DEV)
    x = torch.randn(5, 8, device=DEV)
    assert float(mp.gspmm(ei, None, x, 'sum').abs().sum()) == 0
    assert float(mp.gspmm(ei, None, x, 'max').max()) == -3.4028234663852886e+38
    # shape / rank errors raise (no silent fallback)
    with pytest.raises(IndexError):
        mp.unsorted_segment_sum(torch.ones(3, 2, device=DEV), torch.zeros(3, 1, dtype=torch.int64, device=DEV), 2)
    with pytest.raises(AssertionError):
        mp.unsorted_segment_sum(torch.ones(3, 2, device=DEV), torch.zeros(4, dtype=torch.int64, device=DEV), 2)
    with pytest.raises(Exception):
        mp.gspmm(torch.zeros(2, 3, dtype=torch.int64, device=DEV), None, x, 'add')
    # non-contiguous x, int32 ids, rank-3 x
    xs = torch.randn(40, 6, device=DEV)[:, ::2]
    ids = torch.randint(0, 7, (40,), device=DEV, dtype=torch.int32)
    ref = O.segment('sum', xs.cpu().numpy(), ids.cpu().numpy().astype(np.int64), 7)
    assert np.array_equal(mp.unsorted_segment_sum(xs, ids, 7).cpu().numpy(), ref)
    x3 = torch.randn(40, 3, 5, device=DEV)
    ref = O.segment('mean', x3.cpu().numpy(), ids.cpu().numpy().astype(np.int64), 7)
    assert np.array_equal(mp.unsorted_segment_mean(x3, ids, 7).cpu().numpy(), ref)
    # gspmm has no gradient for the weight (src/gspmm.cpp:30,79)
    ei = torch.tensor([[0, 1, 2], [1, 2, 0]], device=DEV)
    w = torch.ones(3, device=DEV, requires_grad=True)
    xx = torch.randn(3, 4, device=DEV, requires_grad=True)
    mp.gspmm(ei, w, xx).sum().backward()
    assert w.grad is None and xx.grad is not None


def test_determinism_and_cache_identity(mp):
    """Run-to-run identical bits (no float atomics) and the CSR cache must not alias a different graph."""
    N, E, K = 4000, 100000, 100
    index = make_graph(N, E, seed=31, hub=3000).to(DEV)
    x = torch.randn(N, K, device=DEV)
    w = torch.rand(E, device=DEV)
    a = mp.gspmm(index, w, x)
    b = mp.gspmm(index, w, x)
    assert torch.equal(a, b)
    index2 = index.clone()
    index2[1] = (index2[1] + 1) % N
    c = mp.gspmm(index2, w, x)
    ref, _ = O.spmm_fwd('sum', index2.cpu().numpy(), w.cpu().numpy(), x.cpu().numpy())
    assert rel_err(c.cpu().numpy(), ref) < 1e-5
    index.add_(0)                      # in-place op bumps _version -> cache miss, still correct
    assert torch.equal(mp.gspmm(index, w, x), a)


def test_sparse_structures_vs_oracle(mp):
    """gammagl_b200.sparse: ind2ptr / ptr2ind (convert.cpp:58-128), CSRAdj.from_edges (sparse_adj.py:47-67) and the
    COO/CSR/CSC views + matmul of SparseGraph, against numpy / the oracle."""
    from gammagl_b200 import sparse as S
    N, E, K = 700, 9000, 24
    index = make_graph(N, E, seed=41, hub=1500)
    row, col = index[0], index[1]
    srt = np.sort(row.numpy())
    ptr = S.ind2ptr(cu(srt), N)
    assert ptr.dtype == torch.int64 and np.array_equal(ptr.cpu().numpy(), O.ind2ptr(srt, N))
    assert np.array_equal(S.ptr2ind(ptr, E).cpu().numpy(), O.ptr2ind(ptr.cpu().numpy(), E))
    assert np.array_equal(S.ptr2ind(ptr, E).cpu().numpy(), srt)
    adj = S.CSRAdj.from_edges(cu(row), cu(col), N)
    order = np.argsort(row.numpy(), kind='stable')
    assert np.array_equal(adj.degree.cpu().numpy(), np.bincount(row.numpy(), minlength=N))
    assert np.array_equal(adj._sorted_eid.cpu().numpy(), order)
    assert np.array_equal(adj._sorted_u.cpu().numpy(), row.numpy()[order])
    assert np.array_equal(adj._sorted_v.cpu().numpy(), col.numpy()[order])
    assert np.array_equal(adj._indptr.cpu().numpy(), O.ind2ptr(srt, N))
    g = torch.Generator().manual_seed(2)
    h = torch.randn(N, K, generator=g)
    w = torch.rand(E, generator=g)
    # A @ h with A[row, col] = w  ==  gspmm on edge_index (src = col, dst = row)
    ref, _ = O.spmm_fwd('sum', np.stack([col.numpy(), row.numpy()]), w.numpy(), h.numpy())
    got = adj.matmul(cu(h), cu(w)).cpu().numpy()
    short = np.bincount(row.numpy(), minlength=N) <= 1024
    assert np.array_equal(got[short], ref[short]) and rel_err(got, ref) < 1e-5
    sg = S.SparseGraph.from_edge_index(cu(index), cu(w), (N, N))
    assert rel_err(sg.matmul(cu(h)).cpu().numpy(), ref) < 1e-5
    rp, ci, val = sg.csr()
    rc_order = O.sort_edge_index(index.numpy(), N, True)[1]                 # storage.py:124-140: by row * ncols + col
    assert np.array_equal(rp.cpu().numpy(), O.ind2ptr(srt, N)) and np.array_equal(ci.cpu().numpy(), col.numpy()[rc_order])
    assert np.array_equal(val.cpu().numpy(), w.numpy()[rc_order])
    r2, c2, v2 = sg.coo()
    assert np.array_equal(r2.cpu().numpy(), row.numpy()[rc_order]) and np.array_equal(c2.cpu().numpy(), col.numpy()[rc_order])
    cp, ri, _ = sg.csc()
    corder = O.sort_edge_index(index.numpy(), N, False)[1]                  # storage.py:200-217: by col * nrows + row
    assert np.array_equal(ri.cpu().numpy(), row.numpy()[corder])
    # csr2csc maps CSR element order to CSC element order
    assert np.array_equal(val.cpu().numpy()[sg.csr2csc().cpu().numpy()], w.numpy()[corder])
    reft, _ = O.spmm_fwd('sum', index.numpy(), w.numpy(), h.numpy())       # transposed: out[col] += w * h[row]
    assert rel_err(sg.t().matmul(cu(h)).cpu().numpy(), reft) < 1e-5


# ------------------------------------------------------------------------------------------------------
# bipartite blocks and the fused max with the winning edge recorded
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('reduce', ['sum', 'mean', 'max'])
@pytest.mark.parametrize('Ns,Nd,E,K,hub', [(3000, 700, 40000, 100, 2000), (500, 2500, 20000, 32, 0), (900, 900, 15000, 47, 0)])
def test_gspmm_bipartite_rows(mp, reduce, Ns, Nd, E, K, hub):
    """gspmm(..., num_dst=Nd): x holds the Ns source rows, the output the Nd destination rows (SAGEConv's (src, dst) call,
    sage_conv.py:75-92; sampled blocks).  Oracle: message + segment reduce over Nd (what the reference layer computes)."""
    g = torch.Generator().manual_seed(Ns + Nd)
    ei = torch.stack([torch.randint(0, Ns, (E,), generator=g), torch.randint(0, Nd, (E,), generator=g)])
    if hub:
        ei[1, :hub] = 5
    x = torch.randn(Ns, K, generator=g)
    w = torch.rand(E, generator=g)
    gout = torch.randn(Nd, K, generator=g)
    xx = x.to(DEV).requires_grad_(True)
    out = mp.gspmm(ei.to(DEV), w.to(DEV), xx, reduce, num_dst=Nd)
    assert tuple(out.shape) == (Nd, K)
    out.backward(gout.to(DEV))
    msg = (x * 1.0)[ei[0]].numpy() * w.numpy()[:, None]
    short = np.bincount(ei[1].numpy(), minlength=Nd) <= 1024
    got = out.detach().cpu().numpy()
    if reduce == 'max':
        ref, arg = O.segment('max', msg, ei[1].numpy(), Nd)
        assert np.array_equal(got, ref)
        # gspmm-max credits every edge (s -> d) whose SOURCE won (cpu/spmm_max_cpu.cpp:84-96)
        nonempty = np.bincount(ei[1].numpy(), minlength=Nd) > 0
        src_win = np.where(nonempty[:, None], ei[0].numpy()[np.minimum(arg, E - 1)], -1)
        gx = np.zeros((Ns, K))
        s_np, d_np = ei[0].numpy(), ei[1].numpy()
        hit = src_win[d_np] == s_np[:, None]
        np.add.at(gx, s_np, np.where(hit, w.numpy()[:, None].astype(np.float64) * gout.numpy()[d_np], 0.0))
        assert rel_err(xx.grad.cpu().numpy(), gx) < 1e-5
    else:
        ref = O.segment(reduce, msg, ei[1].numpy(), Nd) if reduce == 'sum' else None
        if reduce == 'sum':
            assert np.array_equal(got[short], ref[short])
            assert rel_err(got, ref) < 1e-5
            gm = gout.numpy()[ei[1].numpy()].astype(np.float64)
        else:
            cnt = np.maximum(np.bincount(ei[1].numpy(), minlength=Nd), 1).astype(np.float64)
            sm = np.zeros((Nd, K))
            np.add.at(sm, ei[1].numpy(), msg.astype(np.float64))
            assert rel_err(got, sm / cnt[:, None]) < 1e-5
            gm = (gout.numpy() / cnt[:, None])[ei[1].numpy()]
        gx = np.zeros((Ns, K))
        np.add.at(gx, ei[0].numpy(), gm * w.numpy()[:, None].astype(np.float64))
        assert rel_err(xx.grad.cpu().numpy(), gx) < 1e-5


@pytest.mark.parametrize('weighted', [False, True])
@pytest.mark.parametrize('Ns,Nd,E,K,hub', [(1500, 1500, 30000, 64, 3000), (2000, 300, 25000, 100, 0), (400, 400, 6000, 3, 0),
                                           (300, 900, 8000, 16, 1200)])
def test_gspmm_max_edge_is_message_plus_segment_max(mp, weighted, Ns, Nd, E, K, hub):
    """Fused max with the winning EDGE recorded == `message` + `unsorted_segment_max` of the reference (values, int64 edge
    ids and gradients bit-exact), on a MULTIGRAPH with ties: integer-valued features, duplicated edges."""
    g = torch.Generator().manual_seed(7 * Ns + K)
    ei = torch.stack([torch.randint(0, Ns, (E,), generator=g), torch.randint(0, Nd, (E,), generator=g)])
    ei[:, E // 2:E // 2 + E // 8] = ei[:, :E // 8]                        # parallel edges
    if hub:
        ei[1, :hub] = 2
    x = torch.randint(-4, 5, (Ns, K), generator=g).float()                # ties everywhere
    w = (torch.randint(1, 4, (E,), generator=g).float() * 0.5) if weighted else None
    gout = torch.randn(Nd, K, generator=g)
    xx = x.to(DEV).requires_grad_(True)
    out, arg = mp.gspmm_max_edge(ei.to(DEV), None if w is None else w.to(DEV), xx, num_dst=Nd, return_arg=True)
    out.backward(gout.to(DEV))
    msg = x.numpy()[ei[0].numpy()] * (w.numpy()[:, None] if weighted else 1.0)
    ref, arg_ref = O.segment('max', msg.astype(np.float32), ei[1].numpy(), Nd)
    assert np.array_equal(out.detach().cpu().numpy(), ref)
    nonempty = np.bincount(ei[1].numpy(), minlength=Nd) > 0
    assert np.array_equal(arg.cpu().numpy(), np.where(nonempty[:, None], arg_ref, E))   # empty rows: the oracle's sentinel is N, ours E
    gm = O.segment_bwd('max', gout.numpy(), ei[1].numpy(), E, arg_ref)
    gx = np.zeros((Ns, K), dtype=np.float64)
    np.add.at(gx, ei[0].numpy(), gm.astype(np.float64) * (w.numpy()[:, None] if weighted else 1.0))
    assert rel_err(xx.grad.cpu().numpy(), gx) < 1e-6
