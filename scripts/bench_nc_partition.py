#!/usr/bin/env python
"""BASELINE configs[3]: node classification with full-graph layers on a synthetic AM-shaped graph
(1,666,764 nodes, 266 edge types, 11,976,642 directed edges, D=64, 2 layers, NC default genotype),
destination-partitioned over the GPUs of one node.  One step = forward + cross-entropy over the labelled
nodes + backward + gradient all-reduce + Adam.  Prints one JSON line (rank 0).

  torchrun --nproc-per-node N scripts/bench_nc_partition.py [--scale 0.25] [--steps 5]
"""
import argparse
import json
import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func", defaults=(None,))
NC_GENO = [Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_dense', 2, 1), ('f_sparse', 3, 2), ('f_identity', 4, 3),
                                ('a_sum', 5, 2), ('a_sum', 6, 3), ('a_mean', 7, 4), ('f_dense_last', 8, 7),
                                ('f_sparse_last', 9, 7), ('f_sparse_last', 10, 5)], concat_node=[5, 6, 7, 8, 9, 10]),
           Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_sparse', 2, 1), ('f_identity', 3, 2), ('f_identity', 4, 1),
                                ('a_max', 5, 2), ('a_mean', 6, 3), ('a_mean', 7, 4), ('f_sparse_last', 8, 7),
                                ('f_sparse_last', 9, 8), ('f_identity', 10, 9)], concat_node=[5, 6, 7, 8, 9, 10])]


def run(rank, world, dev, scale=1.0, steps=5, warmup=3, group=None):
    """One measurement of the partitioned NC step on `world` ranks (torch.distributed must be initialised when
    world > 1).  Returns the result dictionary (identical on every rank)."""
    from mr_gnas_b200 import dist as D_
    from mr_gnas_b200.model import Network

    N, ET, E = int(1666764 * scale), 266, int(11976642 * scale)
    D, D0, C, NB, layers, n_lab = 64, 16, 11, 50, 2, 1000
    rng = np.random.RandomState(0)
    perm = rng.permutation(N)
    p = 1.0 / np.power(np.arange(1, N + 1, dtype=np.float64), 0.8)
    p /= p.sum()
    dst = perm[rng.choice(N, size=E, p=p)]
    src = rng.randint(0, N, E)
    et = rng.randint(0, ET, E)
    blocks, part = D_.nc_partition(src, dst, et, N, layers, rank, world, dev, group=group)
    trip_index = torch.from_numpy(np.stack([np.arange(E), src, dst], 1)).to(dev)
    labels = torch.from_numpy(rng.randint(0, C, N)).to(dev)
    idx = torch.from_numpy(rng.choice(N, n_lab, replace=False)).to(dev)
    margs = types.SimpleNamespace(feature_dim=D, op_norm=True)
    torch.manual_seed(0)
    model = Network(dev, NC_GENO, N, C, ET, layers, 1, 2, D, D0, NB, nn.CrossEntropyLoss(), margs).to(dev).train()
    params = list(model.parameters())
    opt = torch.optim.Adam(params, lr=1e-3, fused=True)
    torch.cuda.reset_peak_memory_stats(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = model._loss_partitioned(trip_index, blocks, labels, idx)
        loss.backward()
        D_.allreduce_grads_sum(params, part)
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize()

    loss0 = float(step().detach())      # before any parameter update: must agree across world sizes
    for _ in range(warmup - 1):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    peak = torch.tensor([torch.cuda.max_memory_allocated(dev) / 2 ** 30], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(peak, op=dist.ReduceOp.MAX, group=group)
    out = {"workload": f"c4_am_nc x{scale}: N={N} edge types={ET} E={E} D={D} layers={layers}, "
                       "NC default genotype, full-graph layers, fwd+bwd+Adam (eager)",
           "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": float(ms),
           "mp_edges_per_s": E * layers / (float(ms) / 1e3),
           "edges_local_rank0": part.e_local, "peak_mem_gib_max_rank": float(peak),
           "loss_step0": loss0, "loss": float(loss.detach())}
    del model, opt, blocks, trip_index
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the AM-shaped graph (nodes and edges)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.cuda.set_per_process_memory_fraction(0.92)      # a Python OOM, never a dead box
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = run(rank, world, dev, args.scale, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
