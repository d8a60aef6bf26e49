"""CPU tests (-m "not gpu"): the dense fp64 layer restatements (tests/dense_layers.py) used as layer oracles by the GPU
tests are pinned to outputs AND input gradients of the reference's own layers -- GCNConv (4 norms + weighted), SAGEConv
(mean / gcn / pool), GATConv (3 head configurations), GINConv, SGConv, APPNPConv, executed unmodified on the reference's
compiled CPU operators by tests/golden/make_golden_layers.py (fixture tests/golden/layers_ref.npz)."""
import os

import numpy as np
import pytest
import torch

import dense_layers as D

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'layers_ref.npz'))
CASES = D.cases(Z)


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_fixture_covers_every_layer():
    assert [t for t, _ in CASES] == ['gcn_both', 'gcn_left', 'gcn_right', 'gcn_none', 'gcn_weighted', 'sage_mean', 'sage_gcn',
                                     'sage_pool', 'gat_h8_cat', 'gat_h4_avg', 'gat_h1_cat', 'gin', 'sgc', 'appnp']
    for t, _ in CASES:
        assert t + '_out' in Z.files and t + '_gx' in Z.files and t + '_gout' in Z.files


@pytest.mark.parametrize('tag,fn', CASES, ids=[t for t, _ in CASES])
def test_dense_restatement_matches_reference_layer(tag, fn):
    x = torch.from_numpy(Z['x']).double().requires_grad_(True)
    out = fn(x)
    assert rel(out.detach().numpy(), Z[tag + '_out']) < 2e-6                 # fp32 reference vs fp64 algebra
    out.backward(torch.from_numpy(Z[tag + '_gout']).double())
    assert rel(x.grad.numpy(), Z[tag + '_gx']) < 2e-6
