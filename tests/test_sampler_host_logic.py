"""CPU test of the host logic of gammagl_b200.sampler: the global -> local relabelling (seeds keep their positions, every other
sampled node follows in ascending global id) against a plain-Python restatement of the reference's id map
(/root/reference/gammagl/ops/sparse/cpu/sample.cpp:23-33,47-51), up to the documented permutation of the non-seed nodes."""
import torch

from gammagl_b200 import sampler as S


def test_relabel_maps_back_and_keeps_seed_positions():
    g = torch.Generator().manual_seed(0)
    for n_seeds, n_nbr, N in ((7, 60, 30), (1, 5, 3), (12, 0, 50), (0, 9, 20), (25, 400, 1000)):
        idx = torch.randperm(N, generator=g)[:n_seeds]
        nbr = torch.randint(0, N, (n_nbr,), generator=g)
        n_id, local = S._relabel(idx, nbr)
        assert torch.equal(n_id[:n_seeds], idx)                              # seeds first, in input order
        assert n_id.unique().numel() == n_id.numel()                         # every node once
        assert torch.equal(n_id[local], nbr)                                 # local ids map back to the sampled neighbours
        new = n_id[n_seeds:]
        assert torch.equal(new, torch.sort(new).values)                      # the documented order of the new nodes
        # the reference's map (first-encounter order) holds the same node SET
        seen, order = set(idx.tolist()), list(idx.tolist())
        for v in nbr.tolist():
            if v not in seen:
                seen.add(v)
                order.append(v)
        assert sorted(order) == sorted(n_id.tolist())
