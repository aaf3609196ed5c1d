#!/usr/bin/env python
"""Times mrg_gemm_red (main + fold kernel) at the C1 shapes against torch.mm (cuBLAS fp32), CUDA events, 50 calls."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mr_gnas_b200 import functional as K

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
shapes = [("concat dW", 14541, 200, 800, False), ("linear_e dW", 14541, 200, 200, False), ("DistMult dq", 14541, 256, 200, True),
          ("DistMult dent", 256, 14541, 200, False), ("rel_wt @ emb_e", 475, 475, 200, True)]
for name, rows, F1, F2, km in shapes:
    A = torch.randn(rows, F1, device=dev)
    B = torch.randn(rows, F2, device=dev)
    Ain = A.t().contiguous() if km else A
    fns = {"gemm_red": lambda: K.gemm_red(Ain, B, a_kmajor=km), "torch.mm": lambda: torch.mm(A.t(), B)}
    out = []
    for tag, fn in fns.items():
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()        # replayed from a CUDA graph: device time only, as inside the captured step
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out.append(f"{tag} {e0.elapsed_time(e1) / 100 * 1e3:.1f} us")
    print(f"{name:16s} rows={rows} F1={F1} F2={F2}: " + ", ".join(out))
