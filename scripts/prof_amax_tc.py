#!/usr/bin/env python
"""Where does the tensor pipe of the fused a_max forward idle?  Rebuilds the library with -DMRG_TC_PROF (cycle counters
in the MMA-issuing warp: waiting for the epilogue to drain TMEM, for a W chunk, for an X item, issuing) and runs the
kernel at the C1 shape.  Debugging aid; restores the normal build afterwards."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MRG_TC_PROF"] = "1"
from mr_gnas_b200 import build
build.build(force=True)
import torch, torch.nn as nn
from mr_gnas_b200 import _lib, functional as K
from mr_gnas_b200.graph import MRGraph
from mr_gnas_b200.synth import CONFIGS, synth_kg
lib = _lib.load()
lib.mrg_debug_set_tc_prof.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda:0")
N, R, T, D = CONFIGS["c1_fb15k237"]
g = MRGraph.from_triples(N, synth_kg(N, R, T, seed=0), R, device=dev)
torch.manual_seed(0)
x = torch.relu(torch.randn(g.M, D, device=dev))
lin = nn.Linear(D, D).to(dev)
prof = torch.zeros(32, dtype=torch.int64, device=dev)
for prec in ("fp32", "bf16"):
    K.AMAX_PRECISION = prec
    for _ in range(3):
        K.AMaxTC.apply(x, lin.weight, lin.bias, g, True)
    torch.cuda.synchronize()
    lib.mrg_debug_set_tc_prof(prof.data_ptr())
    prof.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    K.AMaxTC.apply(x, lin.weight, lin.bias, g, True)
    e1.record()
    torch.cuda.synchronize()
    lib.mrg_debug_set_tc_prof(None)
    p = prof.cpu().tolist()
    tot = sum(p[:4])
    print(f"{prec}: call {e0.elapsed_time(e1) * 1e3:.0f} us; MMA warp cycles per item ({p[4]} items over 148 CTAs):")
    for name, v in zip(("wait TMEM drained (epilogue)", "wait W chunk", "wait X item", "issue + rest"), p[:4]):
        print(f"   {name:30s} {100 * v / tot:5.1f} %  {v / max(p[4], 1):8.0f} cycles/item")
    tiles0 = p[4] / (7 if prec == "fp32" else 4) / 2      # tiles of slot 0
    print("   epilogue thread 0 of slot 0, cycles per tile: prelude (edge ids, dst, flags, 3 barriers) %.0f, wait for the "
          "accumulator %.0f, TMEM scan + atomics %.0f (of which tcgen05.ld + wait::ld %.0f)" % (p[5] / tiles0, p[6] / tiles0, p[7] / tiles0, p[8] / tiles0))
    if prec == "fp32" and os.environ.get("MRG_AMAX_PAIR", "1") != "0" and D > 128:
        items_per_cta = p[4] / 74
        for rk in (0, 1):
            v = p[12 + 6 * rk: 18 + 6 * rk]
            print("   producers of CTA rank %d (one thread per group, summed), cycles per item: issue loads %.0f, wait gathered data %.0f, "
                  "wait free X stage %.0f, convert + store %.0f, fence.proxy.async %.0f, syncwarp + arrive %.0f"
                  % ((rk,) + tuple(x / 74 / items_per_cta for x in v)))
os.environ.pop("MRG_TC_PROF")
subprocess.run([sys.executable, "-c", "import os,sys; sys.path.insert(0, %r); from mr_gnas_b200 import build; build.build(force=True)" % ROOT],
               env={k: v for k, v in os.environ.items() if k != "MRG_TC_PROF"})
