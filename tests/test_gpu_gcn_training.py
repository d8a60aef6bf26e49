"""End-to-end plumbing test (BASELINE.json configs[0]): a 2-layer GCN trained on a Cora-shaped graph with the B200
layers follows the same loss trajectory as a dense fp64 restatement of the model (A_hat @ (X @ W) + b) with the same
initial weights and the same Adam steps."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples'))


def test_gcn_two_layer_training_matches_dense_reference():
    import gcn_cora_shaped as ex
    dev = 'cuda:0'
    x, ei, y, train_idx, _ = ex.cora_shaped(dev, seed=3)
    N = x.shape[0]
    torch.manual_seed(3)
    net = ex.GCNModel(x.shape[1], 16, 7, drop_rate=0.0).to(dev)
    # dense fp64 twin
    W1 = net.conv1.linear.weights.detach().double().clone().requires_grad_(True)
    b1 = net.conv1.bias.detach().double().clone().requires_grad_(True)
    W2 = net.conv2.linear.weights.detach().double().clone().requires_grad_(True)
    b2 = net.conv2.bias.detach().double().clone().requires_grad_(True)
    src, dst = ei
    ns = torch.bincount(src, minlength=N).double().pow(-0.5)
    nd = torch.bincount(dst, minlength=N).double().pow(-0.5)
    A = torch.zeros(N, N, dtype=torch.float64, device=dev).index_put_((dst, src), ns[src] * nd[dst], accumulate=True)
    xd = x.double()
    opt = torch.optim.Adam(net.parameters(), lr=0.01, weight_decay=5e-4)
    optd = torch.optim.Adam([W1, b1, W2, b2], lr=0.01, weight_decay=5e-4)
    for epoch in range(30):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(net(x, ei, None, N)[train_idx], y[train_idx])
        loss.backward()
        opt.step()
        optd.zero_grad()
        h = torch.relu(A @ (xd @ W1) + b1)
        lossd = torch.nn.functional.cross_entropy((A @ (h @ W2) + b2)[train_idx], y[train_idx])
        lossd.backward()
        optd.step()
        assert abs(float(loss) - float(lossd)) < 2e-4 * max(1.0, abs(float(lossd))), (epoch, float(loss), float(lossd))
    assert float(loss) < 1.9          # it learns (ln 7 = 1.946 at start)
    assert torch.allclose(net.conv2.linear.weights.detach().double(), W2.detach(), atol=5e-4)
