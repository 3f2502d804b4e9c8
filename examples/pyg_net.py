"""The network of reference models/pytorch_geometric/example.py:27-55 (4 x PNAConvSimple + BatchNorm + residual + mean
pooling + MLP) with the layer class swapped for pna_b200's -- the only change a user of the reference makes.
Runs a few training steps on synthetic molecule-like graphs (there is no network for ogbg-molhiv).

    python examples/pyg_net.py [--steps 20]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.nn import BatchNorm1d, Linear, ModuleList, ReLU, Sequential

from pna_b200 import PNAConvSimple, csr_from_edge_index, synth      # reference: from models.pytorch_geometric.pna import PNAConvSimple


def global_mean_pool(x, batch, n_graphs):
    out = torch.zeros((n_graphs, x.size(1)), dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    cnt = torch.bincount(batch, minlength=n_graphs).clamp(min=1).unsqueeze(1)
    return out / cnt


class Net(torch.nn.Module):
    def __init__(self, deg, hidden=80, n_layers=4):
        super().__init__()
        aggregators, scalers = ["mean", "min", "max", "std"], ["identity", "amplification", "attenuation"]   # example.py:33-34
        self.convs, self.batch_norms = ModuleList(), ModuleList()
        for _ in range(n_layers):
            self.convs.append(PNAConvSimple(hidden, hidden, aggregators, scalers, deg, post_layers=1))
            self.batch_norms.append(BatchNorm1d(hidden))
        self.mlp = Sequential(Linear(hidden, 40), ReLU(), Linear(40, 20), ReLU(), Linear(20, 1))

    def forward(self, x, edge_index, batch, n_graphs):
        csr = csr_from_edge_index(edge_index, x.size(0))          # built once, shared by the 4 layers
        for conv, bn in zip(self.convs, self.batch_norms):
            h = F.relu(bn(conv(x, edge_index, csr=csr)))
            x = F.dropout(h + x, 0.3, training=self.training)     # example.py:50-52
        return self.mlp(global_mean_pool(x, batch, n_graphs))


def main(steps=20, n_graphs=2000, hidden=80, seed=0, verbose=True):
    dev = torch.device("cuda:0")
    torch.manual_seed(seed)
    ei, x, batch = synth.zinc_like(n_graphs=n_graphs, n_feat=hidden, seed=seed)
    y = (torch.bincount(batch, minlength=n_graphs).float().unsqueeze(1) - 23.0) / 4.0       # learnable target: graph size
    deg = synth.degree_histogram(ei[1], x.size(0))
    net = Net(deg, hidden).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    x, ei, batch, y = x.to(dev), ei.to(dev), batch.to(dev), y.to(dev)
    losses = []
    for step in range(steps):
        opt.zero_grad()
        loss = F.mse_loss(net(x, ei, batch, n_graphs), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
        if verbose and step % 5 == 0:
            print(f"step {step:3d} loss {losses[-1]:.4f}")
    return losses


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    main(ap.parse_args().steps)
