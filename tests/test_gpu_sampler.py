"""GPU tests of the neighbour sampler (csrc/sampler.cu, gammagl_b200/sampler.py) against the selection rule of the reference
(/root/reference/gammagl/ops/sparse/cpu/sample.cpp:10-135, neighbor_sample.cpp:22-125):
  * fanout < 0 or >= degree: every neighbour -- the deterministic branch, compared exactly with a numpy restatement;
  * without replacement: exactly min(fanout, degree) DISTINCT positions of the seed's own row; with replacement: fanout draws;
  * uniformity of the draws (each neighbour of a degree-10 node is picked ~ fanout / 10 of the time);
  * multi-hop neighbor_sample: edges connect a frontier node to one of its true in-neighbours, local ids map back through
    `samples`; directed=False returns exactly the induced sub-graph;
  * NeighborSampler blocks feed the fused bipartite aggregation and reproduce message + segment-mean on the block."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def csr_by_dst(ei, N):
    order = np.argsort(ei[1], kind='stable')
    col = ei[0][order]
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(rowptr, ei[1] + 1, 1)
    return np.cumsum(rowptr), col, order


def make(N, E, seed):
    g = np.random.default_rng(seed)
    ei = np.stack([g.integers(0, N, E), g.integers(0, N - N // 10, E)])      # last tenth: no in-edges
    ei[1, :3000] = 4                                                            # a hub
    return ei


def test_sample_adj_all_neighbours_is_exact():
    from gammagl_b200 import sampler as S
    N, E = 2000, 30000
    ei = make(N, E, 1)
    rowptr, col, _ = csr_by_dst(ei, N)
    idx = np.array([4, 7, 1999, 0, 55, 1234], dtype=np.int64)
    for fanout in (-1, 10 ** 6):
        op, oc, n_id, e_id = S.sample_adj(torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV), torch.from_numpy(idx).to(DEV),
                                          fanout, replace=False, seed=3)
        op, oc, n_id, e_id = (t.cpu().numpy() for t in (op, oc, n_id, e_id))
        deg = rowptr[idx + 1] - rowptr[idx]
        assert np.array_equal(op, np.concatenate([[0], np.cumsum(deg)]))
        want_e = np.concatenate([np.arange(rowptr[w], rowptr[w + 1]) for w in idx])
        assert np.array_equal(e_id, want_e)                                    # every neighbour, in CSR order (sample.cpp:37-54)
        assert np.array_equal(n_id[:len(idx)], idx) and np.array_equal(n_id[oc], col[want_e])


@pytest.mark.parametrize('replace', [False, True])
def test_sample_adj_counts_membership_distinctness(replace):
    from gammagl_b200 import sampler as S
    N, E, fanout = 3000, 60000, 5
    ei = make(N, E, 2)
    rowptr, col, _ = csr_by_dst(ei, N)
    idx = np.random.default_rng(0).permutation(N)[:800].astype(np.int64)
    op, oc, n_id, e_id = (t.cpu().numpy() for t in S.sample_adj(torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV),
                                                                 torch.from_numpy(idx).to(DEV), fanout, replace=replace, seed=11))
    deg = rowptr[idx + 1] - rowptr[idx]
    want = np.where(deg == 0, 0, fanout) if replace else np.minimum(deg, fanout)
    assert np.array_equal(np.diff(op), want)
    for i, w in enumerate(idx):
        e = e_id[op[i]:op[i + 1]]
        assert np.all((e >= rowptr[w]) & (e < rowptr[w + 1]))                 # picked from the seed's own row
        if not replace:
            assert len(np.unique(e)) == len(e)                                  # Floyd: no position twice
    assert np.array_equal(n_id[oc], col[e_id])
    # same seed -> same draws; another seed -> other draws
    again = S.sample_adj(torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV), torch.from_numpy(idx).to(DEV), fanout,
                         replace=replace, seed=11)[3].cpu().numpy()
    other = S.sample_adj(torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV), torch.from_numpy(idx).to(DEV), fanout,
                         replace=replace, seed=12)[3].cpu().numpy()
    assert np.array_equal(again, e_id) and not np.array_equal(other, e_id)


def test_sampling_is_uniform():
    from gammagl_b200 import sampler as S
    deg, fanout, trials = 10, 3, 4000
    rowptr = torch.tensor([0, deg], dtype=torch.int64, device=DEV)
    col = torch.arange(deg, dtype=torch.int64, device=DEV)
    # many independent seeds = many copies of the same node in one launch would share the key: vary the seed instead
    counts = np.zeros(deg)
    rp = torch.arange(0, (trials + 1) * deg, deg, dtype=torch.int64, device=DEV)           # `trials` nodes, all of degree 10
    cl = torch.arange(deg, dtype=torch.int64, device=DEV).repeat(trials)
    for replace in (False, True):
        _, _, _, e_id = S.sample_adj(rp, cl, torch.arange(trials, device=DEV), fanout, replace=replace, seed=5)
        c = np.bincount(e_id.cpu().numpy() % deg, minlength=deg)
        expect = trials * fanout / deg
        assert np.all(np.abs(c - expect) < 6 * np.sqrt(expect)), (replace, c)
    assert rowptr.numel() == 2 and col.numel() == deg and counts.sum() == 0


def test_neighbor_sample_multi_hop_and_induced():
    from gammagl_b200 import sampler as S
    N, E = 1500, 20000
    ei = make(N, E, 3)
    colptr, row, _ = csr_by_dst(ei, N)
    seeds = np.array([4, 10, 20, 30], dtype=np.int64)
    samples, rows, cols, edges = (t.cpu().numpy() for t in S.neighbor_sample(
        torch.from_numpy(colptr).to(DEV), torch.from_numpy(row).to(DEV), torch.from_numpy(seeds).to(DEV), [4, 3], replace=False,
        directed=True, seed=1))
    assert np.array_equal(samples[:4], seeds) and len(np.unique(samples)) == len(samples)
    assert np.array_equal(samples[rows], row[edges])                            # local source ids map back to the true sources
    for e, c in zip(edges, cols):                                               # the edge ends in the frontier node it was drawn for
        assert colptr[samples[c]] <= e < colptr[samples[c] + 1]
    samples2, r2, c2, e2 = (t.cpu().numpy() for t in S.neighbor_sample(
        torch.from_numpy(colptr).to(DEV), torch.from_numpy(row).to(DEV), torch.from_numpy(seeds).to(DEV), [4, 3], replace=False,
        directed=False, seed=1))
    assert np.array_equal(samples2, samples)
    inset = np.zeros(N, dtype=bool)
    inset[samples] = True
    dst_of = np.repeat(np.arange(N), np.diff(colptr))
    want = np.nonzero(inset[row] & inset[dst_of])[0]                            # every edge between two sampled nodes
    assert np.array_equal(np.sort(e2), want)
    assert np.array_equal(samples2[r2], row[e2]) and np.array_equal(samples2[c2], dst_of[e2])


def test_neighbor_sampler_blocks_feed_the_bipartite_aggregation():
    from gammagl_b200 import mpops, sampler as S
    N, E, K = 4000, 80000, 32
    ei = make(N, E, 4)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, K, generator=g)
    ns = S.NeighborSampler(torch.from_numpy(ei).to(DEV), [5, 4], num_nodes=N, seed=9)
    batch = torch.arange(100, 164)
    bs, n_id, adjs = ns.sample(batch)
    assert bs == 64 and torch.equal(n_id[:64].cpu(), batch)
    h = x.to(DEV)[n_id]
    for adj in adjs:                                                             # outermost hop first, like the reference loader
        n_src, n_dst = adj.size
        assert h.shape[0] == n_src
        eil = adj.edge_index
        # the sampled edges are edges of the graph: original edge ids map back
        src_g, dst_g = n_id[eil[0]].cpu().numpy(), n_id[eil[1]].cpu().numpy()
        assert np.array_equal(ei[0][adj.e_id.cpu().numpy()], src_g) and np.array_equal(ei[1][adj.e_id.cpu().numpy()], dst_g)
        out = mpops.gspmm(eil, None, h, 'mean', num_dst=n_dst)                    # SAGE-mean on the block, fused
        ref = O.segment('mean', h.cpu().numpy()[eil[0].cpu().numpy()], eil[1].cpu().numpy(), n_dst)
        assert np.max(np.abs(out.cpu().numpy() - ref)) <= 1e-5 * max(1e-30, np.max(np.abs(ref)))
        h = out
    assert h.shape[0] == 64
