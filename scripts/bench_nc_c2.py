#!/usr/bin/env python
"""BASELINE configs[1]: node classification with a fixed searched genotype on a synthetic AIFB-shaped graph
(8,285 nodes, 90 edge types, 58,086 directed edges, D=64, init 16, 2 layers, 4 classes), trained the reference's
way: 64 labelled seeds per step, 2-layer full-neighbour message-flow blocks (train/mr_nc_train.py:42-72).
One step = block sampling (host) is EXCLUDED; forward + cross-entropy + backward + Adam on prebuilt blocks.
Prints one JSON line."""
import json
import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func", defaults=(None,))
NC_GENO = [Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_dense', 2, 1), ('f_sparse', 3, 2), ('f_identity', 4, 3),
                                ('a_sum', 5, 2), ('a_sum', 6, 3), ('a_mean', 7, 4), ('f_dense_last', 8, 7),
                                ('f_sparse_last', 9, 7), ('f_sparse_last', 10, 5)], concat_node=[5, 6, 7, 8, 9, 10]),
           Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_sparse', 2, 1), ('f_identity', 3, 2), ('f_identity', 4, 1),
                                ('a_max', 5, 2), ('a_mean', 6, 3), ('a_mean', 7, 4), ('f_sparse_last', 8, 7),
                                ('f_sparse_last', 9, 8), ('f_identity', 10, 9)], concat_node=[5, 6, 7, 8, 9, 10])]


def main():
    from mr_gnas_b200.graph import full_neighbor_blocks
    from mr_gnas_b200.model import Network
    dev = torch.device("cuda:0")
    N, ET, E, D, D0, C, NB, B = 8285, 90, 58086, 64, 16, 4, 50, 64
    rng = np.random.RandomState(0)
    p = 1.0 / np.power(np.arange(1, N + 1, dtype=np.float64), 0.8)
    p /= p.sum()
    dst = rng.permutation(N)[rng.choice(N, size=E, p=p)]
    src, et = rng.randint(0, N, E), rng.randint(0, ET, E)
    trip_index = torch.from_numpy(np.stack([np.arange(E), src, dst], 1)).to(dev)
    labels = torch.from_numpy(rng.randint(0, C, N)).to(dev)
    train_idx = rng.choice(N, 140, replace=False)
    batches = []
    for k in range(4):
        seeds = np.sort(rng.choice(train_idx, B, replace=False))
        blocks = full_neighbor_blocks(src, dst, et, seeds, 2, device=dev)
        batches.append((blocks, torch.from_numpy(seeds).to(dev), sum(b.E for b in blocks)))
    args = types.SimpleNamespace(feature_dim=D, op_norm=True)
    torch.manual_seed(0)
    model = Network(dev, NC_GENO, N, C, ET, 2, 1, 2, D, D0, NB, nn.CrossEntropyLoss(), args).to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)

    def step(i):
        blocks, seeds, _ = batches[i % len(batches)]
        opt.zero_grad(set_to_none=True)
        loss = model._loss(trip_index, blocks, labels, seeds)
        loss.backward()
        opt.step()
        return loss

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 50
    e0.record()
    for i in range(steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    edges = float(np.mean([b[2] for b in batches]))
    print(json.dumps({"workload": f"c2_aifb_nc: N={N} edge types={ET} E={E} D={D} 2 layers, {B} seeds/step, "
                                  "2-layer full-neighbour blocks, fwd+bwd+Adam (eager, launch bound)",
                      "ms_per_step": ms, "block_edges_per_step": edges, "mp_edges_per_s": edges / (ms / 1e3),
                      "seeds_per_s": B / (ms / 1e3), "loss": float(loss.detach())}), flush=True)


if __name__ == "__main__":
    main()
