"""GPU tests of the error and caching behaviour around the kernels (round-1 advisor findings):
  * ids outside their range raise IndexError (the reference's CPU operators raise as well: TORCH_CHECK_INDEX idx < N,
    /root/reference/gammagl/mpops/torch_ext/cpu/segment_max_cpu.cpp:50) -- nothing is clamped silently;
  * index tensors created under torch.inference_mode() work (no version counter: never cached);
  * two CUDA streams aggregating over the same cached graph do not share the long-row partials buffer;
  * GGLB200_TILE_EDGES above GGLB200_LONG_ROW is refused (a long row must be the last row of its tile)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(scope='module')
def mp():
    from gammagl_b200 import _lib, mpops
    _lib.load()
    return mpops


def test_out_of_range_ids_raise(mp):
    from gammagl_b200 import utils
    from gammagl_b200.csr import clear_caches
    g = torch.Generator().manual_seed(0)
    N, E, K = 50, 400, 8
    x = torch.randn(E, K, generator=g).to(DEV)
    ids = torch.randint(0, N, (E,), generator=g).to(DEV)
    bad = ids.clone()
    bad[17] = N                                    # one past the end
    with pytest.raises(IndexError):
        mp.unsorted_segment_sum(x, bad, N)
    neg = ids.clone()
    neg[3] = -1
    with pytest.raises(IndexError):
        mp.unsorted_segment_max(x, neg, N)
    ei = torch.randint(0, N, (2, E), generator=g).to(DEV)
    xs = torch.randn(N, K, generator=g).to(DEV)
    w = torch.rand(E, generator=g).to(DEV)
    e2 = ei.clone()
    e2[0, 5] = N + 3                               # bad SOURCE id: would be an out-of-bounds gather
    with pytest.raises(IndexError):
        mp.gspmm(e2, w, xs, 'sum')
    e3 = ei.clone()
    e3[1, 5] = N                                   # bad destination id
    with pytest.raises(IndexError):
        mp.gspmm(e3, w, xs, 'sum')
    lg = torch.randn(E, 4, generator=g).to(DEV)
    for path in ('csr', 'oneshot'):                # segment_softmax: the CSR build resp. the id validation of the sort-free form
        old_path, mp._SOFTMAX_PATH = mp._SOFTMAX_PATH, path
        try:
            with pytest.raises(IndexError):
                mp.segment_softmax(lg, bad, N)
        finally:
            mp._SOFTMAX_PATH = old_path
    with pytest.raises(IndexError):
        utils.sort_edge_index(e2, num_nodes=N)
    with pytest.raises(IndexError):
        utils.coalesce(e3, num_nodes=N)
    clear_caches()
    out = mp.gspmm(ei, w, xs, 'sum')               # the library is still usable after the errors
    ref, _ = O.spmm_fwd('sum', ei.cpu().numpy(), w.cpu().numpy(), xs.cpu().numpy())
    assert np.array_equal(out.cpu().numpy(), ref)


def test_inference_mode_index_tensors(mp):
    g = torch.Generator().manual_seed(1)
    N, E, K = 300, 5000, 16
    ei_c = torch.randint(0, N, (2, E), generator=g)
    x_c = torch.randn(N, K, generator=g)
    w_c = torch.rand(E, generator=g)
    ref, _ = O.spmm_fwd('sum', ei_c.numpy(), w_c.numpy(), x_c.numpy())
    with torch.inference_mode():
        ei = ei_c.to(DEV)                          # an inference tensor: ei._version raises
        assert ei.is_inference()
        out = mp.gspmm(ei, w_c.to(DEV), x_c.to(DEV), 'sum')
        seg = mp.unsorted_segment_sum(x_c.to(DEV)[ei[0]], ei[1], N)
        deg = mp.unsorted_segment_sum(torch.ones(E, device=DEV), ei[1], N)
    assert np.array_equal(out.cpu().numpy(), ref)
    r2 = O.segment('sum', x_c.numpy()[ei_c[0].numpy()], ei_c[1].numpy(), N)
    assert np.array_equal(seg.cpu().numpy(), r2)
    assert np.array_equal(deg.cpu().numpy(), np.bincount(ei_c[1].numpy(), minlength=N).astype(np.float32))


def test_two_streams_share_a_cached_graph(mp):
    """Both streams reduce a graph with a split row (> 1024 edges): each must use its own partials buffer."""
    g = torch.Generator().manual_seed(2)
    N, E, K = 200, 30000, 64
    ei = torch.randint(0, N, (2, E), generator=g)
    ei[1, :6000] = 7
    xa, xb = torch.randn(N, K, generator=g), torch.randn(N, K, generator=g)
    w = torch.rand(E, generator=g)
    eid, wd, xad, xbd = ei.to(DEV), w.to(DEV), xa.to(DEV), xb.to(DEV)
    mp.gspmm(eid, wd, xad, 'sum')                  # builds + caches the CSR on the default stream
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(20):
        with torch.cuda.stream(s1):
            oa = mp.gspmm(eid, wd, xad, 'sum')
        with torch.cuda.stream(s2):
            ob = mp.gspmm(eid, wd, xbd, 'sum')
        outs.append((oa, ob))
    torch.cuda.synchronize()
    ra, _ = O.spmm_fwd('sum', ei.numpy(), w.numpy(), xa.numpy())
    rb, _ = O.spmm_fwd('sum', ei.numpy(), w.numpy(), xb.numpy())
    for oa, ob in outs:
        assert np.max(np.abs(oa.cpu().numpy() - ra)) <= 1e-5 * np.max(np.abs(ra))
        assert np.max(np.abs(ob.cpu().numpy() - rb)) <= 1e-5 * np.max(np.abs(rb))
        assert torch.equal(oa, outs[0][0]) and torch.equal(ob, outs[0][1])     # deterministic, no cross-talk


def test_rebuilt_edge_lists_hit_the_cache_by_content(mp):
    """Reference layers build a fresh edge tensor every forward (sage_conv.py:83-85 add_self_loops): an identity miss falls back
    to an exact content match, so the CSR is built once; a tensor modified in place afterwards is never matched."""
    from gammagl_b200 import csr
    g = torch.Generator().manual_seed(4)
    N, E, K = 300, 5000, 12
    ei = torch.stack([torch.randint(0, N, (E,), generator=g), torch.randint(0, N, (E,), generator=g)]).to(DEV)
    x = torch.randn(N, K, generator=g).to(DEV)
    ids = ei[1].clone()
    msg = torch.randn(E, K, generator=g).to(DEV)
    csr.clear_caches()
    h0, s0 = csr._graphs.content_hits, csr._segments.content_hits
    a = mp.gspmm(ei, None, x, 'sum')
    b = mp.gspmm(ei.clone(), None, x, 'sum')                       # new storage, same contents
    assert torch.equal(a, b) and csr._graphs.content_hits == h0 + 1
    sa = mp.unsorted_segment_sum(msg, ids, N)
    sb = mp.unsorted_segment_sum(msg, ids.clone(), N)
    assert torch.equal(sa, sb) and csr._segments.content_hits == s0 + 1
    # in-place change of a cached tensor: its old entry must not serve a clone of the NEW contents
    e2 = ei.clone()
    first = mp.gspmm(e2, None, x, 'sum')
    e2[1, :100] = (e2[1, :100] + 1) % N
    want, _ = O.spmm_fwd('sum', e2.cpu().numpy(), np.ones(E, dtype=np.float32), x.cpu().numpy())
    got = mp.gspmm(e2.clone(), None, x, 'sum')
    assert np.abs(got.cpu().numpy() - want).max() <= 1e-5 * np.abs(want).max()
    assert not torch.equal(got, first)
