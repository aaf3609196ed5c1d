#!/usr/bin/env python
"""Host-side profile (cProfile) of the C3 supernet step at the script's default graph_batch_size (300): the step is
launch/host bound there, this shows where the Python time goes."""
import cProfile, os, pstats, sys, io
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mr_gnas_b200 import _lib
from mr_gnas_b200.model_search_lp import Network
from mr_gnas_b200.synth import CONFIGS, synth_kg
from mr_gnas_b200.utils import weights_init
from mr_gnas_b200.utils_rgcn import sample_search_graph

dev = torch.device("cuda:0")
size = int(sys.argv[1]) if len(sys.argv) > 1 else 300
N, R, T, D = CONFIGS["c3_wn18rr"]
trip_d = torch.from_numpy(synth_kg(N, R, T, seed=0)).to(dev)
torch.manual_seed(0)
model = Network(dev, N, R, 2, 1, 2, 2, D, 100, 2 * R + 1, 40, 0.3, 0.1)
model.apply(weights_init)
model = model.to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
gen = torch.Generator(device=dev).manual_seed(0)

def step():
    s = sample_search_graph(trip_d, size, 0.5, R, 10, device=dev, generator=gen)
    ent, rel = model(s["g"], s["uniq_v"].view(-1, 1), s["src"], s["etype"])
    loss = model.get_loss(s["g"], ent, rel, s["samples"], s["labels"])
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
    opt.step()
    opt.zero_grad(set_to_none=True)

for _ in range(3):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
print("wall per step (with profiler): %.1f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print("wall per step (no profiler): %.1f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
