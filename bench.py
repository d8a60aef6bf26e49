#!/usr/bin/env python
"""bench.py -- edges aggregated / second of the message-passing aggregation path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c3|c2]
                    [--graph rmat|uniform] [--dtype f32|bf16]

Workload (BASELINE.json north_star "Target"): GCN aggregation forward + backward on a synthetic
ogbn-products-shaped graph (C3: N=2,449,029 nodes, E=123,718,280 edges incl. self loops, K=100 fp32),
through the public operator `gspmm(edge_index, weight, x, 'sum')` + autograd backward.  One step = one
forward SpMM + one backward SpMM (transposed CSR) = 2*E edge aggregations.  The CSR is cached per graph
(its build time is reported separately as csr_build_ms).  `--workload c2` is BASELINE.json configs[1]
(N=1M, E=50M, K=128).

N > 1: the graph fits one GPU, so the headline `value` comes from independent replicas on independent graphs (weak
scaling, no data-path collective).  The same JSON line then carries `partitioned`: the 1-D node
partition over the N GPUs through the fused NVLink peer-gather kernel (gammagl_b200.partition.PeerGraph /
gglb200_spmm_csr_peer + the deduplicated halo cache) -- first an in-process parity gate on a small partitioned graph
against the CPU oracle (`parity_ok`), then the SAME C3 graph partitioned over N GPUs (strong scaling) and the per-GPU
shard of the ogbn-papers100M shape (C5, weak scaling), each with ms/step, kernel ms and NVLink GB/s pulled per GPU.
Timing: CUDA events, barrier + synchronize on both sides, max over ranks.  Inputs (x 0.98 GB, CSR 1.5 GB) are far
larger than the 126 MB L2, so no explicit flush.

N = 1 additionally reports `configs`: the other BASELINE.json configurations (C3 on a uniform graph, C2, C3 SAGE-mean,
C4 GAT operators, C3 bf16) with ms, algorithmic bytes and fraction of the measured HBM peak each.

--impl reference: the reference's own CPU operators (oracle/_ref = /root/reference/gammagl/mpops/torch_ext
compiled unmodified; else the C oracle port) on a bounded sample of the same workload, host cores only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N, E incl. self loops, K, description)
    'c3': (2449029, 123718280, 100, 'GCN aggregation fwd+bwd, synthetic ogbn-products shape (2.45M nodes / 123.7M edges / feat 100)'),
    'c2': (1000000, 50000000, 128, 'gspmm fwd+bwd microbench, synthetic 1M nodes / 50M edges / feat 128'),
    # per-GPU shard of the ogbn-papers100M shape (111M nodes / 1.6B edges / feat 128 over 8 GPUs); weak scaling
    'c5part': (13882495, 201960734, 128, 'GCN aggregation fwd+bwd, 1-D node partition + halo all-to-all, ogbn-papers100M shape per-GPU shard (13.9M nodes / 202M edges / feat 128 per GPU)'),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--peer', action='store_true',
                    help='c5part only: fused peer-memory gather (gglb200_spmm_csr_peer) instead of pack + all-to-all')
    ap.add_argument('--graph', default='rmat', choices=['rmat', 'uniform', 'block'],
                    help='block (c5part only): 80%% of the edges have a source on the owning GPU')
    ap.add_argument('--dtype', default='f32', choices=['f32', 'bf16'])
    ap.add_argument('--scale', type=float, default=1.0, help='shrink N and E (debug only; reported in config)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='N=1: skip the extra BASELINE.json configurations')
    ap.add_argument('--no-partitioned', action='store_true', help='N>1: skip the 1-D node partition section')
    ap.add_argument('--no-c5', action='store_true', help='N>1: skip the ogbn-papers100M-shape shards of the partition section')
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------
# synthetic graphs (generated on the device they are used on; seeded)
# ----------------------------------------------------------------------------------------------------
def gen_graph(torch, N, E, kind, seed, device):
    """edge_index[2,E] int64, unsorted: E-N random edges (uniform or RMAT 0.57/0.19/0.19/0.05 with random
    relabelling) + N self loops (as examples/gcn/gcn_trainer.py:58 appends), then a random edge permutation."""
    g = torch.Generator(device=device).manual_seed(seed)
    Er = max(E - N, 0)
    if kind == 'uniform':
        src = torch.randint(0, N, (Er,), generator=g, device=device)
        dst = torch.randint(0, N, (Er,), generator=g, device=device)
    else:
        bits = max(1, (N - 1).bit_length())
        src = torch.zeros(Er, dtype=torch.int64, device=device)
        dst = torch.zeros(Er, dtype=torch.int64, device=device)
        a, b, c = 0.57, 0.19, 0.19
        for _ in range(bits):
            u = torch.rand(Er, generator=g, device=device)
            sbit = (u >= a + b).to(torch.int64)                       # quadrants c, d -> source bit 1
            dbit = (((u >= a) & (u < a + b)) | (u >= a + b + c)).to(torch.int64)   # quadrants b, d -> dest bit 1
            src = (src << 1) | sbit
            dst = (dst << 1) | dbit
            del u, sbit, dbit
        src %= N
        dst %= N
        relabel = torch.randperm(N, generator=g, device=device)
        src = relabel[src]
        dst = relabel[dst]
        del relabel
    loops = torch.arange(N, device=device)
    src = torch.cat([src, loops])[:E]
    dst = torch.cat([dst, loops])[:E]
    perm = torch.randperm(src.numel(), generator=g, device=device)
    ei = torch.stack([src[perm], dst[perm]])
    del src, dst, perm
    return ei


def gcn_norm(torch, ei, N):
    """calc_gcn_norm (gammagl/utils/norm.py:5-30): w = deg^-1/2[src] * deg^-1/2[dst], deg = in-degree by src."""
    deg = torch.bincount(ei[0], minlength=N).to(torch.float32)
    dinv = deg.pow(-0.5)
    dinv[torch.isinf(dinv)] = 0
    return (dinv[ei[0]] * dinv[ei[1]]).contiguous()


class ClockSampler:
    """nvidia-smi clocks/throttle sampler running during the timed region (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.idx = str(gpu_index)
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [s.strip() for s in line.split(',')]
            if len(f) >= 9 and f[0] == self.idx:
                self.rows.append(f)

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(float(r[1])) for r in self.rows if r[1].replace('.', '').isdigit())
        mx = [int(float(r[2])) for r in self.rows if r[2].replace('.', '').isdigit()]
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx[0] if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.rows)}


def ncu_traffic(workload, graph):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full` capture of
    this workload (profiles/ncu_<round>_spmm_ring_<graph>_raw.csv).  The capture is only trusted when its stamp
    (<csv>.stamp.json, written by tools/capture_round.sh next to the CSV) carries the digest of the sources the captured kernel
    is compiled from (gammagl_b200.build.kernel_digest(): spmm_impl.cuh, common.cuh, spmm_f32.cu, the header's tuning constants,
    the compiler flags); otherwise traffic is None and the reason is returned."""
    if workload != 'c3':
        return None, 'no capture for this workload'
    import glob
    from gammagl_b200 import build as _b
    want = _b.kernel_digest()
    cands = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'ncu_r*_spmm_ring_%s_raw.csv' % graph)), reverse=True)
    why = 'no capture under profiles/'
    for path in cands:
        stamp = path + '.stamp.json'
        try:
            st = json.load(open(stamp))
        except Exception:
            why = 'capture %s has no stamp' % os.path.basename(path)
            continue
        if st.get('kernel_digest') != want:
            why = 'stale: %s was captured from other kernel sources' % os.path.basename(path)
            continue
        try:
            import csv
            rows = list(csv.reader(open(path)))
            hdr, units, vals = rows[0], rows[1], rows[2]
            tot = 0.0
            for name in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                i = hdr.index(name)
                v = float(vals[i].replace(',', ''))
                u = units[i].lower()
                tot += v * (1e9 if u.startswith('g') else 1e6 if u.startswith('m') else 1e3 if u.startswith('k') else 1.0)
            return tot, os.path.relpath(path, ROOT)
        except Exception as e:
            why = 'unreadable capture %s: %r' % (os.path.basename(path), e)
    return None, why


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ----------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/)
# ----------------------------------------------------------------------------------------------------
E_CPU_MAX = 20_000_000        # BASELINE.md section 3 / SURVEY.md 8d: E_cpu = min(E, 2e7) at the same average degree and K


def _cpu_info():
    model = None
    try:
        for ln in open('/proc/cpuinfo'):
            if ln.startswith('model name'):
                model = ln.split(':', 1)[1].strip()
                break
    except Exception:
        pass
    return model


def cpu_reference_leg(N, E, K, reps, warmup, graph_kind, measure_message_path=True):
    """Time the reference's own CPU operators on a bounded sample of the workload: E_cpu = min(E, 2e7)
    edges at the same average degree and K, `warmup` untimed + `reps` timed repetitions, median.  Two measurements:
      (2) the fused operator GCNConv / GATConv dispatch to under the torch backend (message_passing.py:144):
          c_spmm_sum forward + backward  -- like for like with the GPU arm, this is `value`;
      (1) the path SAGEConv / GINConv and every layer without message_aggregate run (message_passing.py:56-59,84-86):
          x[src] * w[:, None] -> c_segment_sum, forward and forward + backward.
    The C++ edge loops are single-threaded in the reference build; the ATen ops around them use all host cores."""
    import statistics
    import torch
    from oracle.build_ref import load_ref
    E_cpu = int(min(E, E_CPU_MAX))
    N_cpu = max(2, int(round(N * (E_cpu / E))))
    ei = gen_graph(torch, N_cpu, E_cpu, 'uniform' if graph_kind == 'uniform' else 'rmat', 1234, 'cpu')
    w = gcn_norm(torch, ei, N_cpu)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N_cpu, K, generator=g)
    gout = torch.randn(N_cpu, K, generator=g)
    ref = None
    try:
        ref = load_ref()
    except Exception:
        ref = None

    def timed(fn):
        for _ in range(warmup):
            fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return statistics.median(ts)

    msg_path = None
    if ref is not None:
        kind = 'reference'

        def fused():
            xx = x.clone().requires_grad_(True)
            ref.c_spmm_sum(ei, w, xx).backward(gout)
        dt = timed(fused)
        if measure_message_path:
            # the message tensor and its gradient are [E_m, K] fp32 each (+ the saved gather): bound E_m by the free host memory
            try:
                import psutil
                avail = psutil.virtual_memory().available
            except Exception:
                avail = 32 << 30
            E_m = int(min(E_cpu, max(1_000_000, (avail // 2) // (5 * K * 4))))
            src, dst = ei[0, :E_m].contiguous(), ei[1, :E_m].contiguous()
            wm = w[:E_m].contiguous()

            def msg_fwd():
                with torch.no_grad():
                    return ref.c_segment_sum(x[src] * wm[:, None], dst, N_cpu)

            def msg_fwd_bwd():
                xx = x.clone().requires_grad_(True)
                ref.c_segment_sum(xx[src] * wm[:, None], dst, N_cpu).backward(gout)
            t_f = timed(msg_fwd)
            t_fb = timed(msg_fwd_bwd)
            msg_path = {'op': 'x[src] * w[:, None] -> c_segment_sum (message_passing.py:56-59,84-86)', 'edges': E_m,
                        'fwd_ms': t_f * 1e3, 'fwd_edges_per_s': E_m / t_f, 'fwd_bwd_ms': t_fb * 1e3,
                        'fwd_bwd_edges_per_s': 2 * E_m / t_fb}
    else:
        kind = 'port'
        from oracle import oracle as O
        ein, wn, xn, gn = ei.numpy(), w.numpy(), x.numpy(), gout.numpy()

        def fused():
            O.spmm_fwd('sum', ein, wn, xn)
            O.spmm_bwd('sum', ein, wn, gn)
        dt = timed(fused)
    return {'value': 2 * E_cpu / dt, 'unit': 'edges/s', 'cores': 1, 'kind': kind, 'ms_per_step': dt * 1e3,
            'host_threads_available': os.cpu_count(), 'torch_threads': torch.get_num_threads(), 'cpu_model': _cpu_info(),
            'message_path': msg_path,
            'sample': 'c_spmm_sum fwd+bwd on %d nodes / %d edges (E_cpu = min(E, 2e7), same avg degree and K=%d as the workload), '
                      '%d warm-up + %d timed reps, median; edge loops are single-threaded in the reference build (setup.py:50 never '
                      'enables OpenMP), ATen ops around them use %d threads' % (N_cpu, E_cpu, K, warmup, reps, torch.get_num_threads())}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    try:                                   # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host thread it may use
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        pass
    N, E, K, desc = WORKLOADS[args.workload]
    N, E = int(N * args.scale), int(E * args.scale)
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    r = cpu_reference_leg(N, E, K, steps, warm, args.graph, measure_message_path=True)
    line = {
        'impl': 'reference', 'metric': 'edges aggregated/sec (SpMM+scatter fwd+bwd)', 'value': r['value'], 'unit': 'edges/s',
        'n_gpus': args.gpus, 'steps': steps, 'warmup': warm, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': desc, 'graph': args.graph, 'N': N, 'E': E, 'K': K, 'note': 'CPU, bounded sample per step'},
        'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample', 'cpu_model', 'torch_threads',
                                           'host_threads_available', 'message_path')},
        'e2e': {'value': r['value'], 'unit': 'edges/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# host placement: bind the rank (threads + first-touch pinned buffers) to the NUMA node of its GPU
# ----------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """CPU affinity of this process := the cores of the NUMA node the GPU hangs off (sysfs), BEFORE any pinned allocation, so
    that cudaHostAlloc'ed staging buffers are first-touched on that node and H2D / D2H copies do not cross the socket link.
    Returns a small report for the JSON line."""
    rep = {'bound': False}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        sysfs = '/sys/bus/pci/devices/%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(sysfs + '/numa_node').read().strip())
        rep['numa_node'] = node
        if node < 0:
            return rep
        cpulist = open('/sys/devices/system/node/node%d/cpulist' % node).read().strip()
        cpus = set()
        for part in cpulist.split(','):
            a, _, b = part.partition('-')
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            rep.update(bound=True, cpus=len(allowed))
    except Exception as e:
        rep['error'] = repr(e)[:120]
    return rep


def median_ms(torch, fn, reps=5, warm=2):
    import statistics
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


# ----------------------------------------------------------------------------------------------------
# N = 1: the other BASELINE.json configurations, each with ms, algorithmic bytes and fraction of peak
# ----------------------------------------------------------------------------------------------------
def run_configs(torch, dev, args, peak):
    from gammagl_b200 import _lib, mpops
    from gammagl_b200.csr import _ptr, _stream, clear_caches, graph_of
    lib = _lib.load()
    out = {}
    reps = 5

    def entry(ms, alg_bytes, passes_edges, **kw):
        d = {'ms': ms, 'algorithmic_bytes': int(alg_bytes), 'GBps': alg_bytes / ms / 1e6, 'frac': alg_bytes / ms / 1e6 / peak,
             'edges_per_s': passes_edges / ms * 1e3}
        d.update(kw)
        return d

    def fwd_bwd(ei, w, x, gout, reduce='sum'):
        def step():
            x.grad = None
            mpops.gspmm(ei, w, x, reduce).backward(gout)
        return step

    def spmm_bytes(N, E, K, esz, weighted):
        return E * (K * esz + 4 + (4 if weighted else 0)) + N * (K * esz + 4)

    # ---- C3 on a uniform random graph (no locality: the honest HBM-bound case) + SAGE-mean + bf16 on the RMAT graph ----
    N, E, K, _ = WORKLOADS['c3']
    N, E = int(N * args.scale), int(E * args.scale)
    g = torch.Generator(device=dev).manual_seed(11)
    x = torch.randn(N, K, generator=g, device=dev).requires_grad_(True)
    gout = torch.randn(N, K, generator=g, device=dev)
    ei = gen_graph(torch, N, E, 'uniform', 2000, dev)
    w = gcn_norm(torch, ei, N)
    mpops.PROFILE = []
    ms = median_ms(torch, fwd_bwd(ei, w, x, gout), reps)
    prof, mpops.PROFILE = mpops.PROFILE, None
    kms = [a.elapsed_time(b) for (_, a, b) in prof[-2 * reps:]]
    b1 = spmm_bytes(N, E, K, 4, True)
    out['c3_uniform'] = entry(ms, 2 * b1, 2 * E, op="gspmm(ei, gcn_norm, x, 'sum') fwd+bwd, uniform random graph",
                              kernel_ms=sum(kms) / len(kms), kernel_frac=b1 / (sum(kms) / len(kms)) / 1e6 / peak)
    del ei, w
    clear_caches()
    ei = gen_graph(torch, N, E, 'rmat', 1000, dev)
    ms = median_ms(torch, fwd_bwd(ei, None, x, gout, 'mean'), reps)
    out['c3_sage_mean'] = entry(ms, 2 * spmm_bytes(N, E, K, 4, False), 2 * E,
                                op="gspmm(ei, None, x, 'mean') fwd+bwd (SAGEConv mean aggregation), RMAT graph")
    w = gcn_norm(torch, ei, N)
    xb = x.detach().bfloat16().requires_grad_(True)
    ms = median_ms(torch, fwd_bwd(ei, w, xb, gout.bfloat16()), reps)
    out['c3_bf16'] = entry(ms, 2 * spmm_bytes(N, E, K, 2, True), 2 * E,
                           op="gspmm fwd+bwd, bf16 storage / fp32 accumulate (200-byte rows), RMAT graph")
    del ei, w, x, gout, xb
    clear_caches()
    torch.cuda.empty_cache()

    # ---- C2: unsorted_segment_sum + CSR SpMM microbench (N=1M, E=50M, K=128) ----
    N, E, K, _ = WORKLOADS['c2']
    N, E = int(N * args.scale), int(E * args.scale)
    ei = gen_graph(torch, N, E, 'rmat', 1002, dev)
    w = gcn_norm(torch, ei, N)
    x = torch.randn(N, K, generator=g, device=dev).requires_grad_(True)
    gout = torch.randn(N, K, generator=g, device=dev)
    ms = median_ms(torch, fwd_bwd(ei, w, x, gout), reps)
    out['c2'] = entry(ms, 2 * spmm_bytes(N, E, K, 4, True), 2 * E, op="gspmm fwd+bwd, N=1M / E=50M / K=128, RMAT graph")
    msg = torch.randn(E, K, generator=g, device=dev)
    dst = ei[1].contiguous()
    mpops.unsorted_segment_sum(msg, dst, N)
    ms = median_ms(torch, lambda: mpops.unsorted_segment_sum(msg, dst, N), reps)
    out['c2_segment_sum'] = entry(ms, E * (K * 4 + 8) + N * K * 4, E, op='unsorted_segment_sum(msg[E,128], dst, N) forward (segment CSR cached)')
    del ei, w, x, gout, msg, dst
    clear_caches()
    torch.cuda.empty_cache()

    # ---- C4: GAT 8 heads x 8 channels on the Reddit shape: the operators BASELINE.md asks to report separately ----
    N, E, H, C = int(232965 * args.scale), int(114615892 * args.scale), 8, 8
    K = H * C
    ei = gen_graph(torch, N, E, 'rmat', 1004, dev)
    gr = graph_of(ei, N)
    fwd, bwd = gr.fwd, gr.bwd
    x3 = torch.randn(N, H, C, generator=g, device=dev).requires_grad_(True)
    wh = torch.rand(E, H, generator=g, device=dev)
    x2 = x3.detach().view(N, K)
    bs_b = E * (K * 4 + 4 * H + 4) + N * K * 4
    c4 = {}
    ms = median_ms(torch, lambda: mpops.spmm_csr(fwd, fwd.col, x2, _lib.GGL_SUM, w=wh, wperm=fwd.perm, heads=H), reps)
    c4['bspmm_fwd'] = entry(ms, bs_b, E)
    ms = median_ms(torch, lambda: mpops.spmm_csr(bwd, bwd.col, x2, _lib.GGL_SUM, w=wh, wperm=bwd.perm, heads=H), reps)
    c4['bspmm_bwd_x'] = entry(ms, bs_b, E)
    gw = torch.empty(E, H, device=dev)
    ms = median_ms(torch, lambda: _lib.check(lib.gglb200_bspmm_bwd_w(_ptr(fwd.rowptr), _ptr(fwd.col), _ptr(fwd.perm), _ptr(fwd.plan), _ptr(x2),
                                                                       _ptr(x2), _ptr(gw), N, E, H, C, _stream(dev))), reps)
    c4['bspmm_bwd_w_sddmm'] = entry(ms, bs_b, E)
    dst = ei[1].contiguous()
    lg = wh.clone().requires_grad_(True)
    ms = median_ms(torch, lambda: mpops.segment_softmax(lg, dst, N), reps)
    c4['edge_softmax_fwd'] = entry(ms, E * (2 * 4 * H + 8), E)
    al = mpops.segment_softmax(lg, dst, N)
    ms = median_ms(torch, lambda: al.backward(wh, retain_graph=True), reps)
    c4['edge_softmax_bwd'] = entry(ms, E * (3 * 4 * H + 8), E)
    del al, lg, gw
    ssrc = torch.randn(N, H, generator=g, device=dev).requires_grad_(True)
    sdst = torch.randn(N, H, generator=g, device=dev).requires_grad_(True)
    ms = median_ms(torch, lambda: mpops.gat_aggregate(ei, x3, ssrc, sdst, 0.2), reps)
    c4['gat_fused_fwd'] = entry(ms, bs_b - 4 * H * E, E, op='logits + edge softmax + weighted sum in one kernel (no [E,H] tensor)')
    og = mpops.gat_aggregate(ei, x3, ssrc, sdst, 0.2)
    gy = torch.randn_like(og)
    ms = median_ms(torch, lambda: og.backward(gy, retain_graph=True), reps)
    c4['gat_fused_bwd'] = entry(ms, 2 * (bs_b - 4 * H * E), 2 * E, op='two passes: grad x + grad s_src | grad s_dst')
    # the same layer with the reference's other head width, C = 64 (K = 512, x' = 477 MB: not L2-resident): head groups of 2
    del x3, ssrc, sdst, og, gy
    C2 = 64
    x3 = torch.randn(N, H, C2, generator=g, device=dev).requires_grad_(True)
    ssrc = torch.randn(N, H, generator=g, device=dev).requires_grad_(True)
    sdst = torch.randn(N, H, generator=g, device=dev).requires_grad_(True)
    b64 = E * (H * C2 * 4 + 4) + N * H * C2 * 4
    ms = median_ms(torch, lambda: mpops.gat_aggregate(ei, x3, ssrc, sdst, 0.2), reps)
    c4['gat_fused_fwd_c64'] = entry(ms, b64, E, op='H=8, C=64 (K=512): 4 head-group launches of the fused kernel')
    og = mpops.gat_aggregate(ei, x3, ssrc, sdst, 0.2)
    gy = torch.randn_like(og)
    ms = median_ms(torch, lambda: og.backward(gy, retain_graph=True), reps)
    c4['gat_fused_bwd_c64'] = entry(ms, 2 * b64, 2 * E)
    c4['note'] = "x' (N*K*4 = %.0f MB) is L2-resident: these operators run against the L2 / issue roof, the HBM fraction is reported for continuity" % (N * K * 4 / 1e6)
    out['c4_gat'] = c4
    del ei, x3, wh, x2, og, gy, ssrc, sdst, dst
    clear_caches()
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------
# N > 1: the 1-D node partition over the N GPUs (fused NVLink peer gather + deduplicated halo cache)
# ----------------------------------------------------------------------------------------------------
def partition_parity_gate(torch, dist, dev, world, rank):
    """A small graph, identical on every rank (CPU generator), partitioned over the N GPUs and aggregated through
    PeerGraph / gglb200_spmm_csr_peer; every rank checks ITS rows against the CPU oracle of the whole graph (forward of the plain
    sum bit-exact: same per-row edge order as the serial loop; backward and the GCN-normalised variant at 1e-5 relative) and the
    verdicts are AND-reduced.  The oracle is used here as the checker only (the peer path cannot be tested on one GPU)."""
    import numpy as np
    from gammagl_b200 import partition as P
    from oracle import oracle as O
    N, E, K = 1500 * world + 37, 60000 * world, 100
    g = torch.Generator().manual_seed(4242)
    ei = torch.randint(0, N, (2, E), generator=g)
    ei[0, :4000] = torch.randint(0, 8, (4000,), generator=g)        # hub sources: exercised by the halo cache on every rank
    ei[1, 4000:7000] = N // 2                                          # one split row (> 1024 in-edges)
    x = torch.randn(N, K, generator=g)
    gout = torch.randn(N, K, generator=g)
    lo, hi = P.owned_range(N, world, rank)
    m = (ei[1] >= lo) & (ei[1] < hi)
    graph = P.PeerGraph(ei[:, m].to(dev), N)
    src, dst = ei.numpy()
    ok, worst = True, 0.0
    info = {'cached_rows_fwd': graph.stats_fwd['cached_rows'], 'cached_rows_bwd': graph.stats_bwd['cached_rows']}
    for norm in ('none', 'both'):
        xl = x[lo:hi].to(dev).requires_grad_(True)
        out = P.peer_gcn_propagate(graph, xl, norm)
        out.backward(gout[lo:hi].to(dev))
        torch.cuda.synchronize()
        w = np.ones(E, dtype=np.float32)
        if norm == 'both':
            with np.errstate(divide='ignore'):
                ns = np.bincount(src, minlength=N).astype(np.float64) ** -0.5
                nd = np.bincount(dst, minlength=N).astype(np.float64) ** -0.5
            ns[~np.isfinite(ns)] = 0
            nd[~np.isfinite(nd)] = 0
            w = (ns[src] * nd[dst]).astype(np.float32)
        ref, _ = O.spmm_fwd('sum', ei.numpy(), w, x.numpy())
        gref = O.spmm_bwd('sum', ei.numpy(), w, gout.numpy())
        got, ggot = out.detach().cpu().numpy(), xl.grad.cpu().numpy()
        e1 = float(np.max(np.abs(got - ref[lo:hi])) / np.max(np.abs(ref)))
        e2 = float(np.max(np.abs(ggot - gref[lo:hi])) / np.max(np.abs(gref)))
        worst = max(worst, e1, e2)
        ok = ok and e1 <= 1e-5 and e2 <= 1e-5
        if norm == 'none':
            short = np.bincount(dst, minlength=N)[lo:hi] <= 1024
            ok = ok and bool(np.array_equal(got[short], ref[lo:hi][short]))
    t = torch.tensor([1.0 if ok else 0.0, -worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    info.update(parity_ok=bool(t[0].item() == 1.0), max_rel_err=float(-t[1].item()),
                checked='PeerGraph fwd+bwd, norm none (forward bit-exact on unsplit rows) and both (1e-5), N=%d E=%d K=%d over %d GPUs, '
                        'every rank against the CPU oracle of the whole graph' % (N, E, K, world))
    del graph
    return info


def time_partitioned(torch, dist, dev, world, graph, x, gout, steps, warm):
    """fwd+bwd of peer_gcn_propagate: (ms per step max over ranks, avg kernel ms incl. the halo-cache fill)."""
    from gammagl_b200 import mpops, partition as P

    def step():
        x.grad = None
        P.peer_gcn_propagate(graph, x, 'both').backward(gout)
    for _ in range(warm):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    mpops.PROFILE = []
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    dist.barrier()
    torch.cuda.synchronize()
    prof, mpops.PROFILE = mpops.PROFILE, None
    kms = [p0.elapsed_time(p1) for (_, p0, p1) in prof]
    t = torch.tensor([a.elapsed_time(b) / steps, sum(kms) / max(1, len(kms))], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0].item()), float(t[1].item())


def run_partitioned(torch, dist, dev, world, rank, args, ei_global_fn):
    """`partitioned` object of the N > 1 JSON line."""
    from gammagl_b200 import partition as P
    res = {}
    try:
        res.update(partition_parity_gate(torch, dist, dev, world, rank))
    except Exception as e:
        import traceback
        res.update(parity_ok=False, error=traceback.format_exc()[-800:])
        return res
    steps = max(3, min(args.steps, 10))

    def nvlink(graph, K, ms_kernel):
        byt = torch.tensor([float(graph.nvlink_bytes_per_pass(K, 4, 'x') + graph.nvlink_bytes_per_pass(K, 4, 'g'))], device=dev)
        dist.all_reduce(byt, op=dist.ReduceOp.MAX)
        per_step = float(byt.item())
        return per_step, per_step / (2 * ms_kernel * 1e-3) / 1e9          # bytes pulled per GPU per step, GB/s inside the kernels

    def stats(graph):
        t = torch.tensor([graph.stats_fwd['remote_edges'], graph.stats_fwd['cached_rows'], graph.stats_fwd['direct_remote_edges'],
                          graph.stats_fwd['edges']], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return {'remote_edge_frac': float(t[0] / t[3]), 'halo_cache_rows_total': int(t[1]), 'direct_remote_edges_total': int(t[2])}

    # ---- (1) strong scaling: the SAME C3 graph the replicas use on rank 0, partitioned over the N GPUs ----
    N, E, K, desc = WORKLOADS['c3']
    N, E = int(N * args.scale), int(E * args.scale)
    ei = ei_global_fn()                                            # identical on every rank (same seed)
    lo, hi = P.owned_range(N, world, rank)
    mine = (ei[1] >= lo) & (ei[1] < hi)
    ei_local = ei[:, mine].contiguous()
    del ei, mine
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    graph = P.PeerGraph(ei_local, N)
    torch.cuda.synchronize()
    plan_ms = (time.perf_counter() - t0) * 1e3
    del ei_local
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    x = torch.randn(hi - lo, K, generator=g, device=dev).requires_grad_(True)
    gout = torch.randn(hi - lo, K, generator=g, device=dev)
    ms, kms = time_partitioned(torch, dist, dev, world, graph, x, gout, steps, 3)
    byt, gbps = nvlink(graph, K, kms)
    res['c3_strong'] = {'workload': desc + ', ONE graph partitioned over %d GPUs (strong scaling)' % world, 'graph': args.graph,
                        'value': 2 * E / (ms * 1e-3), 'unit': 'edges/s', 'ms_per_step': ms, 'kernel_ms': kms,
                        'nvlink_bytes_in_per_gpu_per_step': byt, 'nvlink_GBps_in': gbps, 'plan_build_ms': plan_ms, 'steps': steps,
                        'op': "peer_gcn_propagate(graph, x, 'both') fwd+bwd: row scale -> barrier -> halo-cache fill -> fused peer SpMM -> barrier",
                        **stats(graph)}
    del graph, x, gout
    torch.cuda.empty_cache()

    # ---- (2) weak scaling: per-GPU shard of the ogbn-papers100M shape (C5), uniform and 80 %-local graphs ----
    if not args.no_c5:
        n, Eg, K5, desc5 = WORKLOADS['c5part']
        n, Eg = int(n * args.scale), int(Eg * args.scale)
        N5 = n * world
        lo, hi = P.owned_range(N5, world, rank)
        for kind in ('block', 'uniform'):
            g = torch.Generator(device=dev).manual_seed(77 + rank)
            dst = torch.randint(lo, hi, (Eg,), generator=g, device=dev)
            src = torch.randint(0, N5, (Eg,), generator=g, device=dev)
            if kind == 'block':
                loc = torch.rand(Eg, generator=g, device=dev) < 0.8
                src = torch.where(loc, torch.randint(lo, hi, (Eg,), generator=g, device=dev), src)
                del loc
            ei5 = torch.stack([src, dst])
            del src, dst
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            graph = P.PeerGraph(ei5, N5)
            torch.cuda.synchronize()
            plan_ms = (time.perf_counter() - t0) * 1e3
            del ei5
            x = torch.randn(hi - lo, K5, generator=g, device=dev).requires_grad_(True)
            gout = torch.randn(hi - lo, K5, generator=g, device=dev)
            s5 = max(2, min(steps, 4))
            ms, kms = time_partitioned(torch, dist, dev, world, graph, x, gout, s5, 2)
            byt, gbps = nvlink(graph, K5, kms)
            # the model BASELINE.json configs[4] names: GCN, 3 layers, full graph -- Linear (128 -> 128) -> normalised
            # propagate over the partition -> ReLU, three times, forward + backward through autograd
            gcn3 = None
            try:
                from gammagl_b200 import mpops as mpops_
                Ws = [torch.nn.Parameter(torch.randn(K5, K5, generator=g, device=dev) * (K5 ** -0.5)) for _ in range(3)]

                def gcn3_step():
                    x.grad = None
                    for w_ in Ws:
                        w_.grad = None
                    h = x
                    for li, w_ in enumerate(Ws):
                        h = P.peer_gcn_propagate(graph, mpops_.linear_tc(h, w_), 'both')   # tcgen05 Linear (forward + input gradient)
                        if li < 2:
                            h = torch.relu(h)
                    h.backward(gout)
                for _ in range(2):
                    gcn3_step()
                dist.barrier()
                torch.cuda.synchronize()
                a3, b3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a3.record()
                for _ in range(s5):
                    gcn3_step()
                b3.record()
                dist.barrier()
                torch.cuda.synchronize()
                t3 = torch.tensor([a3.elapsed_time(b3) / s5], device=dev)
                dist.all_reduce(t3, op=dist.ReduceOp.MAX)
                gcn3 = {'model': 'GCN, 3 layers, hidden 128 (tcgen05 Linear -> partitioned normalised propagate -> ReLU), forward + backward',
                        'ms_per_step': float(t3.item()), 'edge_passes_per_step': 6,
                        'value': world * 6 * Eg / (float(t3.item()) * 1e-3), 'unit': 'edges/s'}
                del Ws
            except Exception as ex:
                gcn3 = {'error': repr(ex)[:300]}
            res['c5_shard_' + kind] = {'gcn_3_layer': gcn3,
                                       'workload': desc5.replace('halo all-to-all', 'fused peer gather + halo cache') + ' (weak scaling)',
                                       'graph': kind, 'N_total': N5, 'E_total': Eg * world, 'K': K5,
                                       'value': world * 2 * Eg / (ms * 1e-3), 'unit': 'edges/s', 'ms_per_step': ms, 'kernel_ms': kms,
                                       'nvlink_bytes_in_per_gpu_per_step': byt, 'nvlink_GBps_in': gbps, 'plan_build_ms': plan_ms,
                                       'steps': s5, **stats(graph)}
            del graph, x, gout
            torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl b200 needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa = bind_to_gpu_numa_node(local)            # before any pinned allocation: staging buffers land on the GPU's NUMA node
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    from gammagl_b200 import _lib, mpops
    from gammagl_b200.csr import graph_of
    lib = _lib.load()

    N, E, K, desc = WORKLOADS[args.workload]
    N, E = int(N * args.scale), int(E * args.scale)
    dtype = torch.float32 if args.dtype == 'f32' else torch.bfloat16
    esz = 4 if args.dtype == 'f32' else 2

    ei = gen_graph(torch, N, E, args.graph, 1000 + rank, dev)
    w = gcn_norm(torch, ei, N)
    g = torch.Generator(device=dev).manual_seed(rank)
    x = torch.randn(N, K, generator=g, device=dev).to(dtype).requires_grad_(True)
    gout = torch.randn(N, K, generator=g, device=dev).to(dtype)

    # CSR ingest (once per graph; reported separately)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gr = graph_of(ei, N)
    _ = gr.fwd, gr.bwd
    torch.cuda.synchronize()
    csr_ms = (time.perf_counter() - t0) * 1e3          # first build of the process: includes allocator growth and module loading
    from gammagl_b200.csr import Graph
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    g2 = Graph(ei, N)
    _ = g2.fwd, g2.bwd
    torch.cuda.synchronize()
    csr_warm_ms = (time.perf_counter() - t0) * 1e3     # the same two builds again (what a graph that changes per step pays)
    del g2
    max_in = int(gr.fwd.degrees().max())
    max_out = int(gr.bwd.degrees().max())

    def step():
        x.grad = None
        out = mpops.gspmm(ei, w, x, 'sum')
        out.backward(gout)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    mpops.PROFILE = []
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = _lib.launch_count() - l0
    prof, mpops.PROFILE = mpops.PROFILE, None
    clocks = sampler.stop()
    ms = ev0.elapsed_time(ev1) / args.steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * 2 * E / (ms_max * 1e-3)

    # per-launch duration of the dominant kernel group (forward SpMM and backward SpMM are the same kernel)
    kern_ms = [a.elapsed_time(b) for (_, a, b) in prof]
    kavg = sum(kern_ms) / max(1, len(kern_ms))
    alg_bytes = E * (K * esz + 4 + 4) + N * (K * esz + 4)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kavg * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(args.workload, args.graph) if (args.dtype == 'f32' and args.scale == 1.0) else (None, None)
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                'traffic_source': traffic_src,
                'kernel': 'spmm_ring_kernel<float,4,32,1,false> (bulk-copy ring; long-row chunks in the same launch) + spmm_finalize_kernel',
                'algorithmic_bytes_per_launch': alg_bytes, 'avg_launch_ms': kavg, 'launches_timed': len(kern_ms),
                'peak_source': peak_src,
                'formula': 'E*(K*s+4+4) + N*(K*s+4) per SpMM pass (SURVEY.md 8d gather model)'}
    if traffic:
        # the same launch seen from DRAM: bytes that actually crossed HBM (ncu, per launch) over the live launch time.  frac > 1
        # above means gathered rows were served by L2; this figure says how close the launch runs to the memory system itself.
        roofline['dram_achieved'] = traffic / (kavg * 1e-3) / 1e9
        roofline['dram_frac'] = roofline['dram_achieved'] / peak

    # end-to-end through the public operator with HOST buffers (pinned): x and grad_out go up, out and grad_x come back
    e2e = None
    if not args.no_e2e:
        xh = x.detach().cpu().pin_memory()
        gh = gout.cpu().pin_memory()
        oh = torch.empty_like(xh).pin_memory()
        gxh = torch.empty_like(xh).pin_memory()
        # Two copy streams besides the compute stream: uploads (H2D) and downloads (D2H) use different PCIe directions and
        # overlap each other and the kernels; ordering is by events only (no host sync inside the step).  The device-side
        # input buffers are double-buffered, so the upload of step i+1 runs while step i still computes (as a training loop
        # that prefetches its next batch does); every step still uploads its own x / grad_out and downloads out / grad_x.
        up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        ev_x, ev_g, ev_out, ev_gx = (torch.cuda.Event() for _ in range(4))
        bufs = []
        for _ in range(2):
            b = {'xd': torch.empty_like(x.detach()).requires_grad_(True), 'gd': torch.empty_like(gout),
                 'free_x': torch.cuda.Event(), 'free_g': torch.cuda.Event()}
            b['free_x'].record(cur)
            b['free_g'].record(cur)
            bufs.append(b)
        keep = {'n': 0}

        def e2e_step():
            b = bufs[keep['n'] & 1]
            keep['n'] += 1
            xd, gd = b['xd'], b['gd']
            with torch.cuda.stream(up):
                up.wait_event(b['free_x'])                  # the step that last used this buffer no longer reads it
                with torch.no_grad():
                    xd.copy_(xh, non_blocking=True)
                ev_x.record(up)
                up.wait_event(b['free_g'])
                gd.copy_(gh, non_blocking=True)
                ev_g.record(up)
            cur.wait_event(ev_x)
            xd.grad = None
            out = mpops.gspmm(ei, w, xd, 'sum')
            ev_out.record(cur)
            with torch.cuda.stream(down):
                down.wait_event(ev_out)
                oh.copy_(out.detach(), non_blocking=True)
            cur.wait_event(ev_g)
            out.backward(gd)
            b['free_x'].record(cur)
            b['free_g'].record(cur)
            ev_gx.record(cur)
            gx = xd.grad
            with torch.cuda.stream(down):
                down.wait_event(ev_gx)
                gxh.copy_(gx, non_blocking=True)
            out.record_stream(down)
            gx.record_stream(down)
            keep['last'] = (out, gx)

        for _ in range(3):
            e2e_step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            e2e_step()
        cur.wait_stream(down)
        b.record()
        barrier()
        ems = a.elapsed_time(b) / args.steps
        t = torch.tensor([ems], device=dev)
        per_rank = [{'rank': rank, 'ms': ems, 'numa': numa}]
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            gathered = [None] * world
            dist.all_gather_object(gathered, per_rank[0])
            per_rank = gathered
        nb = N * K * esz
        e2e = {'value': world * 2 * E / (float(t.item()) * 1e-3), 'unit': 'edges/s', 'ms_per_step': float(t.item()),
               'h2d_bytes_per_step': 2 * nb, 'd2h_bytes_per_step': 2 * nb,
               'host_GBps_per_gpu_each_direction': 2 * nb / (float(t.item()) * 1e-3) / 1e9,
               'host_memory_GBps_all_ranks': world * 4 * nb / (float(t.item()) * 1e-3) / 1e9, 'numa': numa, 'per_rank': per_rank,
               'note': 'x and grad_out copied from pinned host memory, out and grad_x copied back, every step (uploads and '
                       'downloads on their own streams, event-ordered, device input buffers double-buffered); graph (CSR) resident'}

    # the same GCN aggregation with the normalisation folded into row scalings (mpops.gcn_propagate): reported as an
    # extra, the headline stays on the reference operator gspmm(edge_index, weight, x)
    folded = None
    if args.dtype == 'f32':
        xf = x.detach().clone().requires_grad_(True)

        def fstep():
            xf.grad = None
            mpops.gcn_propagate(ei, xf, 'both').backward(gout)
        for _ in range(3):
            fstep()
        barrier()
        a2, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a2.record()
        for _ in range(min(args.steps, 10)):
            fstep()
        b2.record()
        barrier()
        fms = a2.elapsed_time(b2) / min(args.steps, 10)
        folded = {'op': "gcn_propagate(edge_index, x, 'both') fwd+bwd (row-scale + unweighted SpMM + row-scale)",
                  'ms_per_step': fms, 'value': 2 * E / (fms * 1e-3), 'unit': 'edges/s (this rank)'}
        del xf

    # free the replica workload before the wider sections
    x.grad = None
    del x, gout, w
    if not args.no_e2e:
        del xh, gh, oh, gxh, bufs, keep
    from gammagl_b200.csr import clear_caches
    del gr
    clear_caches()
    torch.cuda.empty_cache()

    configs = None
    if world == 1 and not args.no_configs and args.workload == 'c3' and args.dtype == 'f32':
        configs = run_configs(torch, dev, args, peak)
        if 'c3_uniform' in configs:
            ut, usrc = ncu_traffic('c3', 'uniform') if args.scale == 1.0 else (None, None)
            roofline['uniform_graph'] = {'kernel_ms': configs['c3_uniform']['kernel_ms'], 'frac': configs['c3_uniform']['kernel_frac'],
                                         'traffic': ut, 'traffic_source': usrc,
                                         'dram_frac': (ut / (configs['c3_uniform']['kernel_ms'] * 1e-3) / 1e9 / peak) if ut else None,
                                         'note': 'same kernel and algorithmic bytes on a uniform random graph (no hub rows for the L2 '
                                                 'to serve): the HBM-bound reading of the roofline; the RMAT frac above exceeds 1 '
                                                 'because the gather model charges L2 hits on hub rows to HBM'}
    partitioned = None
    if world > 1 and not args.no_partitioned and args.workload == 'c3' and args.dtype == 'f32':
        del ei
        torch.cuda.empty_cache()
        partitioned = run_partitioned(torch, dist, dev, world, rank, args, lambda: gen_graph(torch, N, E, args.graph, 1000, dev))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_leg(N, E, K, 3, 1, args.graph, measure_message_path=True)
        cpu = {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample', 'cpu_model', 'torch_threads', 'host_threads_available',
                                 'message_path')}

    if rank == 0:
        line = {
            'metric': 'edges aggregated/sec (SpMM+scatter fwd+bwd)', 'value': value, 'unit': 'edges/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_max, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': desc, 'graph': args.graph, 'N': N, 'E': E, 'K': K, 'scale': args.scale,
                       'op': "gspmm(edge_index, gcn_norm_weight, x, 'sum') forward + autograd backward (2 SpMM passes/step)",
                       'l2': 'inputs (x %.2f GB, CSR %.2f GB) exceed the 126 MB L2; no explicit flush' % (N * K * esz / 1e9, E * 12 / 1e9),
                       'parallelism': 'replicas x%d (independent graphs, no collective)' % world,
                       'max_in_degree': max_in, 'max_out_degree': max_out, 'csr_build_ms': csr_ms, 'csr_build_warm_ms': csr_warm_ms},
            'roofline': roofline, 'e2e': e2e, 'cpu_baseline': cpu, 'gpu_launches': int(launches), 'clocks': clocks,
            'folded_gcn_norm': folded, 'configs': configs, 'partitioned': partitioned,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_c5part(args):
    """Partitioned full-graph aggregation (gammagl_b200/partition.py): every rank owns n nodes and the E_g edges that
    point into them; one step = forward (halo all-to-all + SpMM) + backward (SpMM^T + reverse all-to-all + add)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    else:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29531')
        dist.init_process_group('gloo', rank=0, world_size=1)
    from gammagl_b200 import _lib, mpops, partition as P
    _lib.load()
    n, Eg, K, desc = WORKLOADS['c5part']
    n, Eg = int(n * args.scale), int(Eg * args.scale)
    N = n * world
    kind = 'uniform' if args.graph == 'rmat' else args.graph
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    lo, hi = P.owned_range(N, world, rank)
    dst = torch.randint(lo, hi, (Eg,), generator=g, device=dev)
    src = torch.randint(0, N, (Eg,), generator=g, device=dev)
    if kind == 'block':
        loc = torch.rand(Eg, generator=g, device=dev) < 0.8
        src = torch.where(loc, torch.randint(lo, hi, (Eg,), generator=g, device=dev), src)
        del loc
    ei = torch.stack([src, dst])
    del src, dst
    w = torch.rand(Eg, generator=g, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if args.peer:
        plan = P.PeerGraph(ei, N)
        plan.n_halo = 0
    else:
        plan = P.HaloPlan(ei, N)
        plan.csr()
    torch.cuda.synchronize()
    plan_ms = (time.perf_counter() - t0) * 1e3
    del ei
    x = torch.randn(hi - lo, K, generator=g, device=dev).requires_grad_(True)
    gout = torch.randn(hi - lo, K, generator=g, device=dev)

    def step():
        x.grad = None
        if args.peer:
            P.peer_gcn_propagate(plan, x, 'both').backward(gout)
        else:
            P.dist_gspmm_sum(plan, x, w).backward(gout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.15)
    mpops.PROFILE = []
    l0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    launches = _lib.launch_count() - l0
    prof, mpops.PROFILE = mpops.PROFILE, None
    clocks = sampler.stop()
    t = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device=dev)
    halo = torch.tensor([float(plan.n_halo)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(halo, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    kern_ms = [a.elapsed_time(b) for (_, a, b) in prof]
    kavg = sum(kern_ms) / max(1, len(kern_ms))
    alg = Eg * (K * 4 + 8) + (hi - lo) * (K * 4 + 4)
    peak, peak_src = measured_peak()
    if rank == 0:
        line = {
            'metric': 'edges aggregated/sec (SpMM+scatter fwd+bwd)', 'value': world * 2 * Eg / (ms * 1e-3), 'unit': 'edges/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': desc, 'graph': kind, 'N_total': N, 'E_total': Eg * world, 'K': K, 'scale': args.scale,
                       'parallelism': ('1-D node partition x%d, fused peer-memory gather over NVLink (no collective on the data path)' % world)
                       if args.peer else ('1-D node partition x%d, halo all-to-all (NCCL) per pass' % world),
                       'op': 'peer_gcn_propagate (folded GCN norm, unweighted edges)' if args.peer else 'dist_gspmm_sum (per-edge weights)',
                       'halo_rows_per_gpu_max': int(halo.item()), 'halo_bytes_per_pass_per_gpu': int(halo.item()) * K * 4,
                       'plan_build_ms': plan_ms, 'l2': 'inputs exceed the 126 MB L2; no explicit flush'},
            'roofline': {'bound': 'hbm', 'achieved': alg / (kavg * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                         'frac': alg / (kavg * 1e-3) / 1e9 / peak, 'traffic': None,
                         'kernel': 'spmm_ring_kernel' + (' (peer-memory gather: remote rows over NVLink inside the kernel; NVLink-bound on '
                                                         'low-locality graphs, see nvlink_GBps_in)' if args.peer else ' (local SpMM only)'),
                         'algorithmic_bytes_per_launch': alg, 'avg_launch_ms': kavg, 'peak_source': peak_src},
            'exchange': {'ms_per_step_not_in_spmm': ms - 2 * kavg,
                         'note': ('row scaling + 2 device-side barriers per pass; remote rows are fetched inside the SpMM kernel'
                                  if args.peer else 'pack + all_to_all + index_add, forward and backward; not overlapped with the SpMM')},
            'e2e': None, 'cpu_baseline': None, 'gpu_launches': int(launches), 'clocks': clocks,
        }
        print(json.dumps(line))
    dist.destroy_process_group()


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    elif a.workload == 'c5part':
        run_c5part(a)
    else:
        run_b200(a)
