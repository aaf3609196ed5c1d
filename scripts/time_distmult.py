#!/usr/bin/env python
"""Times the fused DistMult 1-N scoring + BCE forward (mrg_distmult_bce_fwd) at the C1 shape, replayed from a CUDA graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mr_gnas_b200 import functional as K
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, N, D = 256, 14541, 200
ent = torch.randn(N, D, device=dev) * 0.3
sub, rel = torch.randn(B, D, device=dev) * 0.3, torch.randn(B, D, device=dev) * 0.3
label = (torch.rand(B, N, device=dev) < 0.01).float() * 0.9 + 0.1 / N
fn = lambda: K.DistMultBCE.apply(ent, sub, rel, label)
for _ in range(5):
    loss = fn()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        fn()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    g.replay()
e1.record(); torch.cuda.synchronize()
z = (sub * rel).double() @ ent.double().t()
ref = torch.nn.functional.binary_cross_entropy_with_logits(z, label.double())
print(f"distmult_bce_fwd B={B} N={N} D={D}: {e0.elapsed_time(e1) / 100 * 1e3:.1f} us per call; loss {float(loss):.8f} vs fp64 {float(ref):.8f}")
