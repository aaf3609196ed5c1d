#!/usr/bin/env python
"""BASELINE configs[2] (C3): LP supernet search step (all candidates mixed by softmax alphas) on the synthetic
WN18RR-shaped KG (N=40,943, R=11, T=86,835; D=200, init 100, num_base_r=23, 2 layers, zero/first/last nodes 1/2/2,
negative_sample 10, split 0.5), at graph_batch_size in {300 (the script default), 30,000, 86,835 (full)}.

One step = what search/mr_lp_search.py:188-236 does per epoch on the weight side: sample the training subgraph
(device pipeline, utils_rgcn.sample_search_graph), supernet forward, BCE-with-logits over the sampled + negative
triplets, backward, gradient clipping, Adam.  Reports the sampling and the model part separately (CUDA events),
one JSON line per batch size.  `--cpu-sampler` times the oracle's numpy restatement of the reference sampler
beside it (the reference's own per-step host cost)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="300,30000,86835")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--cpu-sampler", action="store_true")
    args = ap.parse_args()
    from mr_gnas_b200 import _lib
    from mr_gnas_b200.model_search_lp import Network
    from mr_gnas_b200.synth import CONFIGS, synth_kg
    from mr_gnas_b200.utils import weights_init
    from mr_gnas_b200.utils_rgcn import sample_search_graph
    _lib.load()
    dev = torch.device("cuda:0")
    N, R, T, D = CONFIGS["c3_wn18rr"]
    trip = synth_kg(N, R, T, seed=0)
    trip_d = torch.from_numpy(trip).to(dev)
    torch.manual_seed(0)
    model = Network(dev, N, R, 2, 1, 2, 2, D, 100, 2 * R + 1, 40, 0.3, 0.1)
    model.apply(weights_init)
    model = model.to(dev).train()
    model._device = dev
    for a in model.arch_parameters():
        a.data = a.data.to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
    gen = torch.Generator(device=dev).manual_seed(0)
    for size in [int(s) for s in args.sizes.split(",")]:
        t_samp = t_model = 0.0
        launches0 = None
        for it in range(args.warmup + args.steps):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            if it == args.warmup:
                launches0 = _lib.launch_count
            e0.record()
            s = sample_search_graph(trip_d, size, 0.5, R, 10, device=dev, generator=gen)
            e1.record()
            node_id = s["uniq_v"].view(-1, 1)
            ent, rel = model(s["g"], node_id, s["src"], s["etype"])
            loss = model.get_loss(s["g"], ent, rel, s["samples"], s["labels"])
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
            e2.record()
            torch.cuda.synchronize()
            if it >= args.warmup:
                t_samp += e0.elapsed_time(e1)
                t_model += e1.elapsed_time(e2)
        E = s["g"].E
        out = {"workload": f"c3_wn18rr supernet step: N={N} R={R} T={T} D={D}, graph_batch_size={size}",
               "graph_edges": E, "graph_nodes": s["g"].N, "scored_triplets": int(s["samples"].shape[0]),
               "ms_sampling": t_samp / args.steps, "ms_model": t_model / args.steps,
               "ms_per_step": (t_samp + t_model) / args.steps,
               "mp_edges_per_s": E * 2 / ((t_samp + t_model) / args.steps / 1e3),
               "libmrgnas_launches_per_step": (_lib.launch_count - launches0) / args.steps, "loss": float(loss)}
        if args.cpu_sampler:
            from oracle.mrg_oracle import sample_search_graph as cpu_sampler    # timed beside it, never on the product path
            np.random.seed(0)
            t0 = time.perf_counter()
            cpu_sampler(trip, size, 0.5, R, 10)
            out["ms_sampling_cpu_numpy_port"] = 1e3 * (time.perf_counter() - t0)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
