"""2-layer GCN on a Cora-shaped synthetic graph (BASELINE.json configs[0]: the plumbing config of
/root/reference/examples/gcn/gcn_trainer.py:51-142 -- N=2708, 10556 edges + self loops, 1433 features, hidden 16,
7 classes, Adam lr 0.01, weight decay 5e-4, dropout 0.5).  The Planetoid files are not available offline, so the graph,
features and labels are synthetic (planted 7-block partition); the model and the training loop are the reference's.

    python examples/gcn_cora_shaped.py [--epochs 200]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gammagl_b200.layers import GCNConv, add_self_loops  # noqa: E402


def cora_shaped(device, seed=0, N=2708, E=10556, F=1433, classes=7):
    g = torch.Generator().manual_seed(seed)
    y = torch.randint(0, classes, (N,), generator=g)
    src = torch.randint(0, N, (E,), generator=g)
    same = torch.rand(E, generator=g) < 0.8                      # 80 % of the edges stay inside a class
    cand = torch.randint(0, N, (E, 8), generator=g)
    hit = (y[cand] == y[src][:, None])
    pick = torch.where(hit.any(1), hit.float().argmax(1), torch.zeros(E, dtype=torch.long))
    dst = torch.where(same, cand[torch.arange(E), pick], torch.randint(0, N, (E,), generator=g))
    ei = add_self_loops(torch.stack([src, dst]), N)
    x = (torch.rand(N, F, generator=g) < 0.01).float()
    x[torch.arange(N), y * 7 % F] += 1.0                           # weak class signal
    x = x / x.sum(1, keepdim=True).clamp_min(1)
    idx = torch.randperm(N, generator=g)
    return x.to(device), ei.to(device), y.to(device), idx[:140].to(device), idx[140:640].to(device)


class GCNModel(torch.nn.Module):                                   # gammagl/models/gcn.py:30-64
    def __init__(self, feature_dim, hidden_dim, num_class, drop_rate, fused=True):
        super().__init__()
        self.conv1 = GCNConv(feature_dim, hidden_dim, fused=fused)
        self.conv2 = GCNConv(hidden_dim, num_class, fused=fused)
        self.dropout = torch.nn.Dropout(drop_rate)

    def forward(self, x, edge_index, edge_weight, num_nodes):
        x = torch.relu(self.conv1(x, edge_index, edge_weight, num_nodes))
        x = self.dropout(x)
        return self.conv2(x, edge_index, edge_weight, num_nodes)


def train(epochs=200, drop_rate=0.5, device='cuda:0', seed=0, log=True):
    x, ei, y, train_idx, val_idx = cora_shaped(device, seed)
    torch.manual_seed(seed)
    net = GCNModel(x.shape[1], 16, 7, drop_rate).to(device)
    opt = torch.optim.Adam(net.parameters(), lr=0.01, weight_decay=5e-4)
    losses = []
    for epoch in range(epochs):
        net.train()
        opt.zero_grad()
        logits = net(x, ei, None, x.shape[0])
        loss = torch.nn.functional.cross_entropy(logits[train_idx], y[train_idx])
        loss.backward()
        opt.step()
        losses.append(float(loss))
        if log and (epoch % 20 == 0 or epoch == epochs - 1):
            net.eval()
            with torch.no_grad():
                acc = float((net(x, ei, None, x.shape[0])[val_idx].argmax(1) == y[val_idx]).float().mean())
            print('epoch %3d  train loss %.4f  val acc %.4f' % (epoch, losses[-1], acc))
    return net, losses


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--epochs', type=int, default=200)
    train(ap.parse_args().epochs)
