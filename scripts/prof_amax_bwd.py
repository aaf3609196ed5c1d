#!/usr/bin/env python
"""Phase timing of the a_max backward kernels at C1 (debugging aid): runs the README-genotype step eagerly with
mrg_debug_set_dw_prof armed and prints the dW kernel's per-phase cycle shares, plus CUDA-event times of the
dX-only and dW-only calls."""
import os, sys, types
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mr_gnas_b200 import _lib, functional as K
from mr_gnas_b200._lib import act, ptr
from mr_gnas_b200.graph import MRGraph
from mr_gnas_b200.synth import CONFIGS, synth_kg

dev = torch.device("cuda:0")
N, R, T, D = CONFIGS["c1_fb15k237"]
g = MRGraph.from_triples(N, synth_kg(N, R, T, seed=0), R, device=dev)
E = g.E
torch.manual_seed(0)
x = torch.relu(torch.randn(g.M, D, device=dev))
lin = nn.Linear(D, D).to(dev)
out = K.AMaxTC.apply(x, lin.weight, lin.bias, g, True)
arg = g.last_arg
gout = torch.randn(N, D, device=dev)
lib = _lib.load()
prof = torch.zeros(8, dtype=torch.int64, device=dev)
def run(need_dx, need_dw, n=5):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K.amax_backward(g, gout, arg, act(x), lin.weight, g.M, True, need_dx=need_dx)
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
print("full call (route + dX + dW + fold): %.3f ms" % run(True, True))
print("without dX:                          %.3f ms" % run(False, True))
lib.mrg_debug_set_dw_prof(prof.data_ptr())
prof.zero_()
run(False, True, n=1)
lib.mrg_debug_set_dw_prof(None)
p = prof.cpu().tolist()
tot = sum(p[:5])
print("dW phases (cycles of warp 1 summed over %d CTAs, %d windows): " % (148, p[5]))
for name, v in zip(("in-place activation", "fence+barrier", "publish list", "wait next window", "prefetch+accumulate"), p[:5]):
    print("  %-22s %6.1f %%   %8.0f cycles / window" % (name, 100 * v / tot, v / max(p[5], 1)))
