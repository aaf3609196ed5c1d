#!/usr/bin/env python
"""CUDA-event timing of mrg_amax_bwd at the C1 shape (with and without the dX product): a quick harness for
iterating on the a_max backward kernels outside the full training step."""
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
