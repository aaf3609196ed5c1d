#!/usr/bin/env python
"""bench.py -- MR-GNAS message-passing hot path on B200 (contract: see the task statement).

A "step" is one full training step of the README-genotype link-prediction network on the
synthetic FB15k-237-shaped KG (BASELINE.json configs[0] shape, run on 1xB200): full-graph
message passing through the cell (fwd), DistMult 1-N scoring + BCE, backward, Adam.

  value : MP edges/s = E directed edges x cells / step time, inputs already resident in HBM
  e2e   : same metric through the public API (Network._loss) with the step's batch
          (triples + dense smoothed labels) copied from pinned host memory and the loss read back
  roofline : dominant libmrgnas kernel, algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline : the CPU oracle (port of the reference path) timed on the host cores

--impl reference times the reference's CPU implementation of the same path (the oracle port;
DGL is not installable) on a bounded sample of the workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
import types
from collections import namedtuple

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.nn as nn

Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func")
README_GENOTYPE = [Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_sparse_comp', 2, 1), ('f_sparse_comp', 3, 2),
                                        ('a_max', 4, 2), ('a_max', 5, 3), ('f_sparse_last', 6, 5),
                                        ('f_sparse_last', 7, 5)],
                            concat_node=[4, 5, 6, 7], score_func='sf_DisMult')]
METRIC, UNIT = "mp_edges_per_s_train_step", "edges/s"


def model_args(D):
    return types.SimpleNamespace(feature_dim=D, drop_aggr=0.0, drop_op=0.0, gamma=40, embed_dim=D, conve_hid_drop=0.0,
                                 feat_drop=0.0, num_filt=4, ker_sz=3, k_w=4, k_h=D // 4)


def workload(name):
    from mr_gnas_b200.synth import CONFIGS, synth_kg
    N, R, T, D = CONFIGS[name]
    return N, R, T, D, synth_kg(N, R, T, seed=0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (samples are filtered by their
    timestamps to the [t0, t1] window the caller reports)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self, t0=None, t1=None):
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, sm_all = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk = float(r[1])
                sm_all.append(clk)
                if t0 is not None and not (t0 - 0.02 <= ts <= t1 + 0.02):
                    continue
                sm.append(clk)
                mx.append(float(r[2]))
                for nm, v in zip(names, r[4:8]):
                    if v.strip().lower() == "active":
                        reasons.add(nm)
            except Exception:
                pass
        if not sm and sm_all:   # clock-domain skew between nvidia-smi's timestamps and time.time(): keep the tail
            sm = sm_all[-max(1, len(sm_all) // 3):]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the reference's CPU path for the same step (oracle port -- the reference is Python + the
    un-installable third-party DGL, so nothing of it can travel to the GPU box), all host threads, on the SAME
    full-size workload, batch size, dropout and optimiser as the GPU arm (`--ref-downscale k` samples the first
    T/k triples instead when a quicker run is wanted)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import mrg_oracle as O  # the one place bench.py executes oracle/ as the measured thing
    torch.set_num_threads(os.cpu_count())
    N, R, T, D, trip = workload(args.workload)
    Ts = T if args.ref_downscale <= 1 else max(1000, T // args.ref_downscale)
    out = time_cpu_oracle(O, N, R, trip[:Ts], D, args.batch, args.steps, args.warmup, args.dropout_cell)
    E = 2 * Ts
    ms = out["ms_per_step"]
    val = E * len(README_GENOTYPE) / (ms / 1e3)
    sample = (f"{args.steps} full-size steps (all {T} triples)" if Ts == T else
              f"{args.steps} steps on the first {Ts} of {T} train triples (E={E} directed edges), full N, full D")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "same_config": Ts == T,
            "config": workload_config(args, N, R, T, D, args.batch),
            "run": {"where": f"host CPU, {torch.get_num_threads()} threads", "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "triples_per_s": args.batch / (ms / 1e3), "loss": out["loss"]}
    print(json.dumps(line), flush=True)


def workload_config(args, N, R, T, D, B, parallelism="none (one device)"):
    """The `config` object: identical for the GPU arm and the reference arm at N=1 (run details go to `run`)."""
    return {"workload": f"{args.workload}: README genotype LP train step (fwd+bwd+Adam), "
                        f"N={N} R={R} T={T} E={2 * T} D={D} B={B}, 1 cell",
            "dropout_cell": args.dropout_cell, "drop_aggr": "n/a (no a_sum in the README genotype)", "lbl_smooth": 0.1,
            "optimizer": "Adam lr=1e-3", "parallelism": parallelism}


def oracle_state(N, R, D, seed=0):
    """Seeded initial state of the README-genotype network as a functional parameter dict for the oracle."""
    from mr_gnas_b200.model_lp import Network
    from mr_gnas_b200.utils import weights_init
    torch.manual_seed(seed)
    m = Network('cpu', README_GENOTYPE, N, R, D, D, 2 * R + 1, nn.BCELoss(), 0.0, model_args(D))
    m.apply(weights_init)
    return {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in m.state_dict().items()}


def time_cpu_oracle(O, N, R, trip, D, B, steps, warmup, dropout_cell=0.0, first_batch=None):
    """`steps` timed Adam steps of the oracle.  `first_batch` = (subj, rel, labels) used for step 0 (the parity
    probe: same init, same batch as the GPU arm's probe); -> ms/step, last loss, step-0 loss and gradients."""
    graph = O.build_graph(N, trip, R)
    P = oracle_state(N, R, D)
    params = [v for v in P.values() if v.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-3)
    items = None
    times, first = [], None
    for it in range(warmup + steps):
        if it == 0 and first_batch is not None:
            subj, rel, labels = first_batch
        else:
            if items is None:
                items = O.process_1n(trip, R)
            chunk = items[(it * B) % max(1, len(items) - B):][:B]
            subj = torch.tensor([c["triple"][0] for c in chunk])
            rel = torch.tensor([c["triple"][1] for c in chunk])
            labels = O.smoothed_labels(chunk, N, 0.1)
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = O.bce_loss(O.network_lp(README_GENOTYPE, P, graph, subj, rel, R, training=True,
                                       dropout_cell=dropout_cell), labels)
        loss.backward()
        if it == 0:
            first = (float(loss.detach()), {k: v.grad.detach().clone() for k, v in P.items() if v.grad is not None})
        opt.step()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return {"ms_per_step": 1e3 * sum(times) / len(times), "loss": float(loss.detach()), "first": first}


# ------------------------------------------------------------------------------------------
def algo_bytes(key, M, E, N, D):
    """Compulsory HBM bytes of one call (inputs read once + outputs written once, fp32 rows of
    b = 4*D bytes; index / gate vectors included; L2-resident [N,D] tables excluded)."""
    b = 4 * D
    name = key.split("(")[0]
    rows = int(key.split("(")[1].split(",")[0]) if "(" in key and key.split("(")[1][0].isdigit() else M
    table = {
        "mrg_compose_fwd": rows * (b + 8),                 # write y, read 2 int32 indices (h, r tables L2-resident)
        "mrg_affine_act": rows * 2 * b,                    # read y, write s
        "mrg_colstats": rows * b,
        "mrg_bn_bwd_reduce": rows * 2 * b,                 # read ds, y
        "mrg_bn_bwd_apply": rows * 3 * b,                  # read ds, y; write dy
        "mrg_sparse_gate_fwd": rows * (2 * b + 8),         # read x (x is xin here or 3b), write y, gate+scale
        "mrg_sparse_gate_bwd": rows * (3 * b + 8),         # read dy, x; write dx
        "mrg_seg_reduce_bwd": rows * b,                    # write dm (g/arg tables L2-resident)
    }
    return table.get(name)


# C-ABI call -> the kernels it launches that move HBM data (ncu names, template arguments dropped)
CALL_KERNELS = {
    "mrg_amax_bwd": ["amax_bwd_dw_kernel", "amax_bwd_dx_kernel", "amax_route_kernel"],
    "mrg_amax_tc_fwd": ["tc::amax_tc2_kernel", "tc::amax_tc_kernel"],
    "mrg_sparse_gate_bwd_fused": ["gate_bwd_pipe_kernel"],
    "mrg_sparse_gate_fwd": ["sparse_gate_fwd_kernel"],
    "mrg_bn_bwd_apply": ["bn_bwd_apply_kernel"],
    "mrg_bn_bwd_reduce": ["bn_bwd_reduce_kernel"],
    "mrg_compose_fwd": ["compose_fwd_kernel"],
    "mrg_seg_reduce_fwd": ["seg_reduce_chunk_kernel"],
}


def ncu_traffic(call_key):
    """(DRAM bytes (read + write) per launch of the kernels behind `call_key`, source file) from the newest
    committed `ncu --set full` capture of the same workload (profiles/rNN_traffic.json); (None, None) if absent."""
    cands = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))
    if not cands:
        return None, None
    src = "profiles/" + cands[-1]
    cap = json.load(open(os.path.join(ROOT, src)))
    total, found = 0.0, False
    for k in CALL_KERNELS.get(call_key.split("(")[0], []):
        hits = [v["dram_bytes_per_launch"] for name, v in cap.items() if name.split("<")[0].endswith(k.split("::")[-1])]
        if hits:
            total += max(hits)      # the full-size launch (small node-level launches share the kernel name)
            found = True
            if call_key.startswith("mrg_amax_tc_fwd"):
                break               # one main kernel per call: the CTA pair when captured, else the single-CTA kernel
    return (total if found else None), src


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1_fb15k237")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--dropout-cell", type=float, default=0.3,
                    help="dropout after each cell (train/mr_lp_train.py default 0.3); the parity probe always runs at 0")
    ap.add_argument("--ref-downscale", type=int, default=1,
                    help="--impl reference: time the first T/k triples instead of the full workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-json", default=None, help="write the per-call CUDA-event profile here")
    ap.add_argument("--sparse-labels", action="store_true",
                    help="e2e leg: send the batch's object lists (CSR) and expand the smoothed [B,N] labels on the "
                         "device (mrg_labels_from_csr) instead of copying the dense matrix over PCIe")
    ap.add_argument("--kernels-only", action="store_true",
                    help="profiling aid (ncu): eager warm-up + steps only, no clock sampling / e2e / CPU legs, no JSON")
    ap.add_argument("--no-graph", action="store_true", help="issue every step eagerly instead of replaying the CUDA graph")
    ap.add_argument("--mode", default="auto", choices=["auto", "partition", "dp"],
                    help="N>1: 'partition' (default) = the north-star split: destinations 1-D partitioned over the "
                         "GPUs, NCCL halo all-gather, global BatchNorm statistics, entity-sharded 1-N scoring, ONE "
                         "shared query batch; 'dp' = replicated full-graph MP with per-rank query batches")
    ap.add_argument("--partition", action="store_true", help="same as --mode partition")
    ap.add_argument("--dp", action="store_true", help="same as --mode dp")
    ap.add_argument("--strong", action="store_true",
                    help="partition mode: keep the C1 graph at every N (strong scaling) instead of growing the "
                         "triples with N (weak scaling, the default: per-GPU edge work stays that of C1)")
    ap.add_argument("--amax-bf16", action="store_true",
                    help="run the fused a_max forward in its bf16 variant (mrg_amax_tc_fwd_bf16; tolerance class 2e-2, "
                         "not the fp32 parity path) -- reported as dtype 'f32 + bf16 a_max operands'")
    ap.add_argument("--no-bf16-variant", action="store_true", help="skip the bf16_variant sub-record")
    ap.add_argument("--no-c4", action="store_true", help="skip the AM-shaped NC partition sub-record (c4_partition)")
    ap.add_argument("--c4-scale", type=float, default=1.0)
    ap.add_argument("--c4-steps", type=int, default=3)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from mr_gnas_b200 import _lib
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.model_lp import Network
    from mr_gnas_b200.process_data import make_batch, make_batch_sparse, train_items
    from mr_gnas_b200.synth import CONFIGS, synth_kg
    from mr_gnas_b200.utils import weights_init
    _lib.load()  # fails loudly if the CUDA extension is missing
    if args.amax_bf16:
        from mr_gnas_b200 import functional as _K
        _K.AMAX_PRECISION = "bf16"

    mode = "dp" if args.dp else ("partition" if args.partition else args.mode)
    part_mode = world > 1 and mode in ("auto", "partition")
    grow = world if (part_mode and not args.strong) else 1
    N0, R, T0, D = CONFIGS[args.workload]
    # weak scaling: `grow` x the C1 triples over the C1 entity set (same generator law), so every GPU's edge-level
    # work is that of C1 while the replicated entity table does not grow with the GPU count
    N, T = N0, T0 * grow
    trip = synth_kg(N, R, T, seed=0)
    E, M, B = 2 * T, 2 * T + N, args.batch
    torch.cuda.set_per_process_memory_fraction(0.92)       # a Python OOM, never a dead box
    if part_mode:
        from mr_gnas_b200.dist import lp_partition
        g = lp_partition(trip, N, R, rank, world, device=dev)
    else:
        g = MRGraph.from_triples(N, trip, R, device=dev)
    torch.manual_seed(0)
    model = Network(dev, README_GENOTYPE, N, R, D, D, 2 * R + 1, nn.BCELoss(), args.dropout_cell, model_args(D))
    model.apply(weights_init)
    model = model.to(dev).train()
    params = [p for p in model.parameters()]
    opt = torch.optim.Adam(params, lr=1e-3, fused=True, capturable=True)
    nb = args.steps + args.warmup
    rng = np.random.RandomState(100 + (0 if (part_mode or world == 1) else rank))
    n_queries = train_items(trip, R, select=[])[1]
    n_host_batches = 2 if N * B * 4 > (1 << 28) else 8      # dense [B, N] label rows are pinned on the host
    sels = [rng.choice(n_queries, size=B, replace=False) for _ in range(min(nb, n_host_batches))]
    flat_items, _ = train_items(trip, R, select=np.concatenate(sels))
    host_batches, sparse_batches = [], []
    for i in range(len(sels)):
        its = flat_items[i * B:(i + 1) * B]
        t_h, y_h = make_batch(its, N, lbl_smooth=0.1, pin=True)
        sparse_batches.append(make_batch_sparse(its, pin=True))
        if part_mode:   # this rank scores its own entity range only: it needs its columns of the label matrix
            y_h = y_h[:, g.part.lo:g.part.hi].contiguous().pin_memory()
        host_batches.append((t_h, y_h))
    dev_batches = [(t.to(dev), y.to(dev)) for t, y in host_batches]
    h2d = host_batches[0][0].numel() * 8 + host_batches[0][1].numel() * 4

    # ---- parity probe (1 GPU): loss and every parameter gradient of ONE step from the seeded init on batch 0,
    # dropout 0; the cpu_baseline leg below runs the oracle on exactly this init and batch and reports the match
    probe = None
    if world == 1 and not args.no_cpu_baseline and not args.kernels_only:
        state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        model._dropout = 0.0
        t_d, y_d = dev_batches[0]
        loss_p = model._loss(g, t_d[:, 0], t_d[:, 1], y_d)
        loss_p.backward()
        probe = (float(loss_p), {k: p.grad.detach().cpu().clone() for k, p in model.named_parameters()
                                 if p.grad is not None},
                 (host_batches[0][0][:, 0].clone(), host_batches[0][0][:, 1].clone(), host_batches[0][1].clone()))
        # a live loss tensor keeps its autograd graph -- and with it every parameter's AccumulateGrad node, which
        # remembers the (legacy default) stream it was created on -- alive into the CUDA-graph capture below
        del loss_p
        model.zero_grad(set_to_none=True)
        model.load_state_dict(state0)          # BatchNorm buffers back to the init
        model._dropout = args.dropout_cell

    from mr_gnas_b200.dist import allreduce_grads as _allreduce, allreduce_grads_sum as _allreduce_sum

    def allreduce_grads():
        if part_mode:
            _allreduce_sum(params, g.part)
        else:
            _allreduce(params, world)

    from mr_gnas_b200.train import GraphedTrainStep
    label_cols = host_batches[0][1].shape[1]
    sparse_cfg = None
    if args.sparse_labels:
        sparse_cfg = dict(num_ent=N, lbl_smooth=0.1, cap=2 * max(b[2].numel() for b in sparse_batches),
                          col_lo=g.part.lo if part_mode else 0, col_hi=g.part.hi if part_mode else N)
        h2d = sparse_batches[0][0][:, :2].numel() * 8 + sparse_batches[0][1].numel() * 4 + \
            max(b[2].numel() for b in sparse_batches) * 4
    runner = GraphedTrainStep(model, g, opt, B, label_cols, grad_sync=lambda ps: allreduce_grads(),
                              sparse_labels=sparse_cfg)
    if args.sparse_labels:      # static CSR inputs must hold a valid batch before warm-up / capture
        t0_, p0_, i0_ = sparse_batches[0]
        runner.load_sparse(t0_[:, 0], t0_[:, 1], p0_, i0_)

    def step_eager(i):                       # per-call profiling pass and --no-graph
        trip_d, y_d = dev_batches[i % len(dev_batches)]
        opt.zero_grad(set_to_none=True)
        loss = model._loss(g, trip_d[:, 0], trip_d[:, 1], y_d)
        loss.backward()
        allreduce_grads()
        opt.step()
        return loss

    dev_sparse = [(t.to(dev), p_.to(dev), i_.to(dev)) for t, p_, i_ in sparse_batches] if args.sparse_labels else None

    def step_resident(i):                    # inputs already in HBM
        if args.sparse_labels:
            t_d, p_d, i_d = dev_sparse[i % len(dev_sparse)]
            return runner(t_d[:, 0], t_d[:, 1], label_csr=(p_d, i_d))
        trip_d, y_d = dev_batches[i % len(dev_batches)]
        return runner(trip_d[:, 0], trip_d[:, 1], y_d)

    primed = {"ok": False}

    def step_e2e(i):                         # inputs in pinned host memory, loss read back
        # Every step's batch crosses PCIe inside the timed region.  The copy of batch i+1 is enqueued on a side
        # stream right after step i is enqueued, so it overlaps the step (double-buffered staging set) the way a
        # DataLoader-fed loop would; step i itself consumes the batch staged during step i-1.
        if args.sparse_labels:               # kilobytes per step: plain copies, nothing to overlap
            t_h, p_h, i_h = sparse_batches[i % len(sparse_batches)]
            return runner(t_h[:, 0], t_h[:, 1], label_csr=(p_h, i_h)).item()
        if not primed["ok"]:
            t_h, y_h = host_batches[i % len(host_batches)]
            runner.prefetch(t_h[:, 0], t_h[:, 1], y_h)
            primed["ok"] = True
        loss = runner()
        t_n, y_n = host_batches[(i + 1) % len(host_batches)]
        runner.prefetch(t_n[:, 0], t_n[:, 1], y_n)
        return loss.item()                   # device -> host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sample_clocks=False):
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sample_clocks:                # let nvidia-smi start up while the GPU stays under load; EVERY rank runs the
            for j in range(40):          # same number of steps (each step holds a collective when world > 1)
                fn(j)
            torch.cuda.synchronize()
        barrier()
        t0 = time.time()
        k_before = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            out = fn(i)
        e1.record()
        barrier()
        t1 = time.time()
        timed.launches = _lib.launch_count - k_before
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, (sampler.stop(t0, t1) if sampler else None), out

    k0 = _lib.launch_count
    step_eager(0)
    lib_launches_per_step = _lib.launch_count - k0
    if args.kernels_only:
        for i in range(args.warmup + args.steps):
            step_eager(i)
        torch.cuda.synchronize()
        return
    if not args.no_graph:
        if not args.sparse_labels:
            trip_d, y_d = dev_batches[0]
            runner.load(trip_d[:, 0], trip_d[:, 1], y_d)
        runner.capture()
    for i in range(args.warmup):
        step_resident(i)
    ms, clocks, last_loss = timed(step_resident, args.steps, sample_clocks=True)
    launches = lib_launches_per_step * args.steps   # libmrgnas kernels executed inside the timed region
    for i in range(2):
        step_e2e(i)
    ms_e2e, _, last_e2e = timed(step_e2e, args.steps)

    cells = len(README_GENOTYPE)
    # MP edges are counted ONCE per step whatever the mode: in the partition the ranks share one pass over the
    # (grown) graph; in dp every rank repeats the same full-graph pass -- only the query batches differ
    value = E * cells / (ms / 1e3)
    e2e_value = E * cells / (ms_e2e / 1e3)
    q_units = 1 if (part_mode or world == 1) else world      # 1-N queries (train "triples") per step = q_units * B

    # per-call CUDA-event profile of one step (rank 0) -> dominant kernel + roofline
    roofline, prof_rows = None, []
    if rank == 0:
        _lib.start_profile()
    for i in range(3):                   # all ranks step together (gradient all-reduce inside); rank 0 records
        step_eager(i)
    barrier()
    if rank == 0:
        prof = _lib.stop_profile()
        tot = sum(v[1] for v in prof.values()) / 3
        hbm, how = peaks()
        for key, (cnt, t, nb) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
            ab = (nb / cnt) if nb else algo_bytes(key, g.M, g.E, g.N, D)
            avg_ms = t / cnt
            prof_rows.append({"call": key, "launches_per_step": cnt / 3, "avg_ms": avg_ms, "ms_per_step": t / 3,
                              "share_of_lib_time": t / 3 / tot if tot else None,
                              "algo_gbs": (ab / avg_ms / 1e6) if ab else None, "algo_bytes": ab})
        top = next((r for r in prof_rows if r["algo_gbs"]), None)
        if top:
            traffic, tsrc = ncu_traffic(top["call"])
            roofline = {"bound": "hbm", "kernel": top["call"], "achieved": top["algo_gbs"], "peak": hbm, "unit": "GB/s",
                        "frac": top["algo_gbs"] / hbm, "traffic": traffic,
                        "traffic_source": f"{tsrc} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the "
                                          "call's kernels, per launch)" if tsrc else None,
                        "algorithmic_bytes": top["algo_bytes"], "peak_source": how,
                        "lib_ms_per_step": tot, "step_ms": ms}
        if args.profile_json:
            json.dump(prof_rows, open(args.profile_json, "w"), indent=1)

    cpu_baseline = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import mrg_oracle as O  # cpu_baseline leg only: the checker, timed beside the product
        torch.set_num_threads(os.cpu_count())
        out = time_cpu_oracle(O, N, R, trip, D, B, steps=1, warmup=0, dropout_cell=0.0, first_batch=probe[2])
        cpu_baseline = {"value": E * cells / (out["ms_per_step"] / 1e3), "unit": UNIT, "cores": torch.get_num_threads(),
                        "kind": "port", "sample": "1 full training step (no warm-up) on the full workload: the same "
                                                  "seeded init and query batch as the GPU parity probe, dropout 0",
                        "ms_per_step": out["ms_per_step"]}
        loss_o, grads_o = out["first"]
        loss_g, grads_g, _ = probe
        gmax = max(float(v.norm()) for v in grads_o.values())
        rows = []
        for k, go in grads_o.items():
            n_o = float(go.double().norm())
            err = float((grads_g[k].double() - go.double()).norm()) / max(n_o, 1e-300)
            rows.append((err, k, n_o))
        live = [r for r in rows if r[2] > 1e-4 * gmax]      # (a bias feeding a BatchNorm has an exactly-zero true grad)
        worst = max(live)
        parity = {"loss_gpu": loss_g, "loss_oracle": loss_o, "loss_rel": abs(loss_g - loss_o) / abs(loss_o),
                  "worst_grad_rel": worst[0], "worst_grad": worst[1], "grad_tensors_compared": len(live),
                  "grad_tensors_skipped_zero_true_gradient": sorted(r[1] for r in rows if r[2] <= 1e-4 * gmax),
                  "metric": "||g_gpu - g_oracle||_2 / ||g_oracle||_2 per parameter tensor, same seeded init and batch "
                            "as GPU step 0, dropout 0, fp32 both sides",
                  "note": "fp32-vs-fp32: the real reference's own fp32 error against an fp64 run at this shape is "
                          "~1e-5 and its BCELoss is discontinuous where sigmoid saturates (4 % of the logits at "
                          "this init); tests/test_gpu_config_parity.py holds every tensor to max(1e-5, 4 x the "
                          "reference's own error) against the fp64 truth"}

    # ---- the same step with the bf16 variant of the fused a_max forward: its own CUDA graph,
    # same model / optimiser / batches, timed like `value`; reported beside the fp32 number, never instead of it
    bf16_variant = None
    if not args.amax_bf16 and not args.no_graph and not args.sparse_labels and not args.no_bf16_variant:
        from mr_gnas_b200 import functional as _K
        _K.AMAX_PRECISION = "bf16"
        try:
            runner16 = GraphedTrainStep(model, g, opt, B, label_cols, grad_sync=lambda ps: allreduce_grads())
            trip_d, y_d = dev_batches[0]
            runner16.load(trip_d[:, 0], trip_d[:, 1], y_d)
            runner16.capture()

            def step16(i):
                t_d, yy = dev_batches[i % len(dev_batches)]
                return runner16(t_d[:, 0], t_d[:, 1], yy)
            for i in range(args.warmup):
                step16(i)
            ms16, _, _ = timed(step16, args.steps)
            bf16_variant = {"what": "same step, fused a_max forward with bf16 operands (mrg_amax_tc_fwd_bf16; stated "
                                    "tolerance 2e-2 on the a_max output, tests/test_gpu_ops_lp.py::test_amax_bf16_variant)",
                            "ms_per_step": ms16, "value": E * cells / (ms16 / 1e3), "unit": UNIT}
            del runner16
        finally:
            _K.AMAX_PRECISION = "fp32"

    c4 = None
    if not args.no_c4:
        # BASELINE configs[3]: AM-shaped NC full-graph layers, destination-partitioned over these same N ranks
        # (strong scaling of a fixed graph); measured after the LP step's buffers are released
        del runner
        torch.cuda.empty_cache()
        try:
            sys.path.insert(0, os.path.join(ROOT, "scripts"))
            import bench_nc_partition
            c4 = bench_nc_partition.run(rank, world, dev, scale=args.c4_scale, steps=args.c4_steps, warmup=3)
        except Exception as ex:            # never lose the main line to the sub-record
            c4 = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    if rank == 0:
        scal = "strong scaling on the named workload" if args.strong else \
            f"weak scaling: {grow}x the C1 triples over the C1 entities (per-GPU edge work = C1's)"
        par = "none (one device)" if world == 1 else (
            f"dst-partition x{world}: 1-D destination ranges balanced by in-edges, NCCL halo all-gather, global "
            f"BatchNorm statistics, entity-sharded 1-N scoring, one shared query batch; {scal}" if part_mode else
            f"dp{world} over query batches (replicated full-graph MP per rank, NCCL grad all-reduce; MP edges "
            "counted once)")
        cfg = workload_config(args, N, R, T, D, B, par)
        run = {"where": f"{world}xB200",
               "l2": f"edge tensors are {g.M * D * 4 / 1e6:.0f} MB each (> 126 MB L2); no explicit flush",
               "launch": "eager" if args.no_graph else "whole step replayed from one CUDA graph",
               "labels": ("object lists (CSR) sent per step, expanded + smoothed on the device"
                          if args.sparse_labels else "dense smoothed [B,N] fp32 matrix per step")}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong" if (part_mode and args.strong) else "weak", "vs_baseline": None,
                "dtype": "f32 + bf16 a_max operands" if args.amax_bf16 else "f32",
                "data": "synthetic", "config": cfg, "run": run,
                "triples_per_s": q_units * B / (ms / 1e3),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e, "triples_per_s": q_units * B / (ms_e2e / 1e3)},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "parity": parity, "bf16_variant": bf16_variant, "c4_partition": c4, "loss": float(last_loss)}
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL communicators referenced by a live CUDA graph can stall destroy_process_group(): leave together,
        # after everything is printed, without tearing the communicator down
        sys.stdout.flush()
        dist.barrier()
        torch.cuda.synchronize()
        os._exit(0)


if __name__ == "__main__":
    main()
