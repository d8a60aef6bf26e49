"""Per-operator microbenchmarks on one B200 (CUDA events, 5 warm-ups + 10 reps, median): time, edges/s and achieved
algorithmic GB/s (SURVEY.md 8d formulas) for every kernel on the path.  Output: one JSON object per line.
Usage: python tools/microbench.py [c2|c3|c4]"""
import json
import statistics
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from gammagl_b200 import _lib, mpops  # noqa: E402
from gammagl_b200.csr import build_csr, graph_of  # noqa: E402

dev = torch.device('cuda', 0)
PEAK = bench.measured_peak()[0]


def timeit(fn, reps=10, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def report(name, ms, E, bytes_alg, **kw):
    print(json.dumps(dict(op=name, ms=round(ms, 4), edges_per_s=E / ms * 1e3, alg_GBps=bytes_alg / ms / 1e6,
                          frac_of_measured_hbm=bytes_alg / ms / 1e6 / PEAK, **kw)), flush=True)


def main(cfg):
    if cfg == 'c4':      # Reddit shape, GAT 8 heads x 8 channels
        N, E, H, C = 232965, 114615892, 8, 8
    elif cfg == 'c2':
        N, E, H, C = 1000000, 50000000, 8, 16
    else:
        N, E, H, C = 2449029, 123718280, 4, 25
    K = H * C
    ei = bench.gen_graph(torch, N, E, 'rmat', 7, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(N, K, generator=g, device=dev)
    w = torch.rand(E, generator=g, device=dev)
    wh = torch.rand(E, H, generator=g, device=dev)
    gr = graph_of(ei, N)
    s = 4
    spmm_b = E * (K * s + 8) + N * (K * s + 4)
    report('csr_build(dst)', timeit(lambda: build_csr(ei[1], ei[0], N), reps=5, warm=2), E, E * 16 * 4, cfg=cfg)
    from gammagl_b200 import utils as gutils
    report('sort_edge_index (two-key radix sort, int64 in/out)', timeit(lambda: gutils.sort_edge_index(ei, num_nodes=N), reps=5, warm=2), E, E * (16 + 16 + 5 * 24), cfg=cfg)
    report('coalesce (sort + duplicate groups + unique)', timeit(lambda: gutils.coalesce(ei, num_nodes=N), reps=5, warm=2), E, E * (16 + 16 + 5 * 24 + 24), cfg=cfg)
    fwd = gr.fwd
    bwd = gr.bwd
    for red, code in (('sum', _lib.GGL_SUM), ('mean', _lib.GGL_MEAN), ('max', _lib.GGL_MAX)):
        report('gspmm_%s fwd K=%d' % (red, K), timeit(lambda: mpops.spmm_csr(fwd, fwd.col, x, code, w=w, wperm=fwd.perm,
               want_arg=(red == 'max'))), E, spmm_b + (N * K * 8 if red == 'max' else 0), cfg=cfg)
    report('gspmm_sum bwd_x (transposed CSR)', timeit(lambda: mpops.spmm_csr(bwd, bwd.col, x, _lib.GGL_SUM, w=w, wperm=bwd.perm)), E, spmm_b, cfg=cfg)
    report('gspmm_sum fwd unweighted', timeit(lambda: mpops.spmm_csr(fwd, fwd.col, x, _lib.GGL_SUM)), E, spmm_b - 4 * E, cfg=cfg)
    xb = x.bfloat16()
    report('gspmm_sum fwd bf16', timeit(lambda: mpops.spmm_csr(fwd, fwd.col, xb, _lib.GGL_SUM, w=w, wperm=fwd.perm)), E,
           E * (K * 2 + 8) + N * (K * 2 + 4), cfg=cfg)
    # bspmm family (GAT)
    bs_b = E * (K * s + 4 * H + 4) + N * K * s
    report('bspmm fwd H=%d C=%d' % (H, C), timeit(lambda: mpops.spmm_csr(fwd, fwd.col, x, _lib.GGL_SUM, w=wh, wperm=fwd.perm, heads=H)), E, bs_b, cfg=cfg)
    report('bspmm bwd_x', timeit(lambda: mpops.spmm_csr(bwd, bwd.col, x, _lib.GGL_SUM, w=wh, wperm=bwd.perm, heads=H)), E, bs_b, cfg=cfg)
    lib = _lib.load()
    from gammagl_b200.csr import _ptr, _stream
    gw = torch.empty(E, H, device=dev)
    report('bspmm bwd_w (SDDMM)', timeit(lambda: _lib.check(lib.gglb200_bspmm_bwd_w(_ptr(fwd.rowptr), _ptr(fwd.col), _ptr(fwd.perm), _ptr(fwd.plan), _ptr(x), _ptr(x), _ptr(gw), N, E, H, C, _stream(dev)))), E, bs_b, cfg=cfg)
    dst_ids = ei[1].contiguous()
    lg = wh.clone().requires_grad_(True)
    report('segment_softmax fwd H=%d (max+exp+sum+div)' % H, timeit(lambda: mpops.segment_softmax(lg, dst_ids, N)), E, E * (2 * 4 * H + 8), cfg=cfg)
    al = mpops.segment_softmax(lg, dst_ids, N)
    report('segment_softmax bwd', timeit(lambda: al.backward(wh, retain_graph=True)), E, E * (3 * 4 * H + 8), cfg=cfg)
    if mpops.gat_fusable(H, C):
        xs3 = x.view(N, H, C).clone().requires_grad_(True)
        ssrc = torch.randn(N, H, generator=g, device=dev).requires_grad_(True)
        sdst = torch.randn(N, H, generator=g, device=dev).requires_grad_(True)
        report('gat_aggregate fwd (logits+softmax+bspmm fused)', timeit(lambda: mpops.gat_aggregate(ei, xs3, ssrc, sdst, 0.2)), E, bs_b - 4 * H * E, cfg=cfg)
        og = mpops.gat_aggregate(ei, xs3, ssrc, sdst, 0.2)
        gy = torch.randn_like(og)
        report('gat_aggregate bwd (2 passes: grad x, s_src | grad s_dst)', timeit(lambda: og.backward(gy, retain_graph=True)), E, 2 * (bs_b - 4 * H * E), cfg=cfg)
        # the composition it replaces, end to end through the unfused operators
        def unfused():
            e = torch.nn.functional.leaky_relu(ssrc[ei[0]] + sdst[ei[1]], 0.2)
            return mpops.bspmm(ei, mpops.segment_softmax(e, ei[1], N), xs3)
        report('  unfused: gathers + leaky_relu + segment_softmax + bspmm fwd', timeit(unfused, reps=5, warm=2), E, bs_b, cfg=cfg)
        ou = unfused()
        report('  unfused: backward of the same', timeit(lambda: ou.backward(gy, retain_graph=True), reps=5, warm=2), E, 2 * bs_b, cfg=cfg)
    # segment ops on a materialised message (the reference's unfused path); K limited so that [E,K] fits comfortably
    Ks = min(K, 32)
    msg = torch.randn(E, Ks, generator=g, device=dev)
    dst = ei[1].contiguous()
    mpops.unsorted_segment_sum(msg, dst, N)  # builds + caches the segment CSR
    seg_b = E * (Ks * s + 8) + N * Ks * s
    for red, f in (('sum', mpops.unsorted_segment_sum), ('mean', mpops.unsorted_segment_mean), ('max', mpops.unsorted_segment_max)):
        report('unsorted_segment_%s fwd K=%d' % (red, Ks), timeit(lambda: f(msg, dst, N)), E, seg_b, cfg=cfg)
    go = torch.randn(N, Ks, generator=g, device=dev)
    m2 = msg.clone().requires_grad_(True)
    o = mpops.unsorted_segment_sum(m2, dst, N)
    def seg_bwd():                      # grad set to None first, as optimizer.zero_grad(set_to_none=True) does: otherwise torch ADDS the new
        m2.grad = None                  # [E,K] gradient to the old one (3 extra passes over 14.7 GB -- the "10.9 ms" of the round-1 table)
        o.backward(go, retain_graph=True)
    report('unsorted_segment_sum bwd through autograd (grad reset per step)', timeit(seg_bwd), E, seg_b, cfg=cfg)
    from gammagl_b200.csr import segment_csr_of
    sc = segment_csr_of(dst, N)
    gxb = torch.empty(E, Ks, device=dev)
    report('segment_sum bwd kernel: row gather (gglb200_segment_gather_bwd)', timeit(lambda: _lib.check(lib.gglb200_segment_gather_bwd(
        _lib.GGL_SUM, _ptr(go), _ptr(dst), _ptr(sc.rowptr), _ptr(gxb), E, Ks, _lib.GGL_F32, _stream(dev)))), E, seg_b, cfg=cfg)
    report('segment_sum bwd kernel: scatter through the CSR (gglb200_segment_expand_bwd)', timeit(lambda: _lib.check(lib.gglb200_segment_expand_bwd(
        _lib.GGL_SUM, _ptr(go), _ptr(sc.rowptr), _ptr(sc.perm), _ptr(sc.plan), _ptr(gxb), E, N, Ks, _lib.GGL_F32, _stream(dev)))), E, seg_b, cfg=cfg)
    report('segment_mean bwd kernel: row gather', timeit(lambda: _lib.check(lib.gglb200_segment_gather_bwd(
        _lib.GGL_MEAN, _ptr(go), _ptr(dst), _ptr(sc.rowptr), _ptr(gxb), E, Ks, _lib.GGL_F32, _stream(dev)))), E, seg_b, cfg=cfg)
    report('segment_mean bwd kernel: scatter through the CSR', timeit(lambda: _lib.check(lib.gglb200_segment_expand_bwd(
        _lib.GGL_MEAN, _ptr(go), _ptr(sc.rowptr), _ptr(sc.perm), _ptr(sc.plan), _ptr(gxb), E, N, Ks, _lib.GGL_F32, _stream(dev)))), E, seg_b, cfg=cfg)
    ones = torch.ones(E, device=dev)
    report('degree: unsorted_segment_sum(ones[E]) K=1 (policy auto: sort-free one-shot pass)', timeit(lambda: mpops.unsorted_segment_sum(ones, dst, N)), E, E * 12 + N * 4, cfg=cfg)
    # the sort-free path on wider rows (what ids that change every step pay instead of sort + CSR pass)
    for red, code in (('sum', _lib.GGL_SUM), ('mean', _lib.GGL_MEAN), ('max', _lib.GGL_MAX)):
        report('segment_%s one-shot (gglb200_segment_reduce) K=%d' % (red, Ks), timeit(lambda: mpops.segment_reduce_oneshot(msg, dst, N, code)), E,
               seg_b + (N * Ks * 8 if red == 'max' else 0), cfg=cfg)
    m4 = torch.randn(E, 4, generator=g, device=dev)
    report('segment_sum one-shot K=4', timeit(lambda: mpops.segment_reduce_oneshot(m4, dst, N, _lib.GGL_SUM)), E, E * 24 + N * 16, cfg=cfg)


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'c3')
