"""Mini-batch GraphSAGE on the GPU sampler (the reference's examples/graphsage/reddit_sage.py pattern: NeighborSampler blocks,
SAGEConv on bipartite (x_src, x_dst) pairs), synthetic products-shaped graph -- data sets are not available offline.

    python examples/sage_minibatch.py [--nodes 200000] [--edges 8000000] [--batch 1024] [--steps 30]

Every block is aggregated by the fused bipartite kernel (mpops.gspmm(..., num_dst=n_dst)): no [E, K] message tensor."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gammagl_b200.layers import SAGEConv  # noqa: E402
from gammagl_b200.sampler import NeighborSampler  # noqa: E402


class SAGE(torch.nn.Module):                      # gammagl/models/graphsage.py: two SAGEConv layers over the sampled blocks
    def __init__(self, f_in, hidden, classes, fuse_gather=False):
        super().__init__()
        self.fuse_gather = fuse_gather
        self.convs = torch.nn.ModuleList([SAGEConv(f_in, hidden, aggr='mean', aggregate_first=fuse_gather),
                                          SAGEConv(hidden, classes, aggr='mean')])

    def forward(self, x, adjs, n_id=None):
        """fuse_gather: `x` is the FULL feature matrix; the first block is aggregated straight from it through global source ids
        (n_id[src_local]) and the neighbour Linear runs on the n_dst aggregated rows -- no x[n_id] gather of the n_src sampled rows,
        no Linear over them.  Otherwise `x` = x_full[n_id] as in the reference example."""
        for i, adj in enumerate(adjs):
            n_src, n_dst = adj.size
            if i == 0 and self.fuse_gather:
                ei = torch.stack([n_id[adj.edge_index[0]], adj.edge_index[1]])
                x = self.convs[0]((x, x[n_id[:n_dst]]), ei)
            else:
                x = self.convs[i]((x, x[:n_dst]), adj.edge_index)
            if i + 1 < len(adjs):
                x = torch.relu(x)
        return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nodes', type=int, default=200000)
    ap.add_argument('--edges', type=int, default=8000000)
    ap.add_argument('--feat', type=int, default=100)
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--fuse-gather', action='store_true', help='aggregate the first block straight from the full feature matrix')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    g = torch.Generator(device=dev).manual_seed(0)
    ei = torch.randint(0, a.nodes, (2, a.edges), generator=g, device=dev)
    y = torch.randint(0, 47, (a.nodes,), generator=g, device=dev)
    x = torch.randn(a.nodes, a.feat, generator=g, device=dev) + torch.nn.functional.one_hot(y, 47).float() @ torch.randn(47, a.feat, generator=g, device=dev)
    sampler = NeighborSampler(ei, [25, 10], num_nodes=a.nodes, seed=0)
    net = SAGE(a.feat, 256, 47, fuse_gather=a.fuse_gather).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=3e-3)
    perm = torch.randperm(a.nodes, generator=g, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for step in range(a.steps):
        batch = perm[(step * a.batch) % (a.nodes - a.batch):][:a.batch]
        bs, n_id, adjs = sampler.sample(batch)
        opt.zero_grad(set_to_none=True)
        out = net(x, adjs, n_id) if a.fuse_gather else net(x[n_id], adjs)
        loss = torch.nn.functional.cross_entropy(out, y[batch])
        loss.backward()
        opt.step()
        if step % 10 == 0:
            print('step %3d  loss %.4f  sampled nodes %d  edges %s' % (step, float(loss), n_id.numel(), [int(ad.edge_index.shape[1]) for ad in adjs]))
    torch.cuda.synchronize()
    print('%.1f ms / step (sample + forward + backward + Adam)' % ((time.perf_counter() - t0) / a.steps * 1e3))


if __name__ == '__main__':
    main()
