#!/usr/bin/env python
"""Times the fused a_max forward at the C1 shape (CUDA events, 20 calls after 5 warm-ups) and checks it against the
single-CTA kernel's result when run with MRG_AMAX_PAIR=0 in a second process (the switch is read once per process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn as nn
from mr_gnas_b200 import functional as K
from mr_gnas_b200.graph import MRGraph
from mr_gnas_b200.synth import CONFIGS, synth_kg

def run():
    dev = torch.device("cuda:0")
    N, R, T, D = CONFIGS["c1_fb15k237"]
    D = int(os.environ.get("AMAX_D", D))
    g = MRGraph.from_triples(N, synth_kg(N, R, T, seed=0), R, device=dev)
    torch.manual_seed(0)
    x = torch.relu(torch.randn(g.M, D, device=dev))
    lin = nn.Linear(D, D).to(dev)
    for prec in ("fp32", "bf16"):
        K.AMAX_PRECISION = prec
        for _ in range(5):
            out = K.AMaxTC.apply(x, lin.weight, lin.bias, g, True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            out = K.AMaxTC.apply(x, lin.weight, lin.bias, g, True)
        e1.record()
        torch.cuda.synchronize()
        print(f"pair={os.environ.get('MRG_AMAX_PAIR', '1')} {prec} D={D}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call "
              f"(3 launches); checksum {out.double().sum().item():.6f} absmax {out.abs().max().item():.6f}")
        if prec == "fp32":
            torch.save(out.cpu(), f"/tmp/amax_out_pair{os.environ.get('MRG_AMAX_PAIR', '1')}.pt")

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        run()
    else:
        for pair in ("1", "0"):
            env = dict(os.environ, MRG_AMAX_PAIR=pair)
            subprocess.run([sys.executable, __file__, "child"], env=env, timeout=300)
        a, b = torch.load("/tmp/amax_out_pair1.pt"), torch.load("/tmp/amax_out_pair0.pt")
        print(f"pair vs single-CTA: max|diff| {float((a - b).abs().max()):.3e}, bit-identical {bool((a == b).all())}")
