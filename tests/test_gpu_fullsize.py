"""Parity at BASELINE.json's FULL C1 size (N=14,541, R=237, T=272,115, E=544,230, D=200), where the CPU oracle
needs minutes per step: size-independent properties of the domain plus a same-size recomputation with plain torch
fp32 ops on the GPU (the checker, never the product path).

  * graph: every edge appears once in the dst-CSR / src-CSC / relation segments, ids ascending inside a segment
  * a_sum: column sums of the output == column sums of all edge-expanded rows (sum is permutation invariant)
  * a_max: fused tcgen05 kernel == relu(x W^T + b) -> per-destination max (torch scatter_reduce), argmax is an
    in-edge of its destination that attains the maximum, isolated destinations -> 0 / -1
  * gather backward: deterministic CSC / relation reductions == index_add_ in fp64
  * cell: the fused edge chain == the module-by-module path on the same parameters (loss and all gradients);
    two runs are bit-identical (determinism); the loss falls over 5 Adam steps
"""
import types
from collections import namedtuple

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func")
README = [Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_sparse_comp', 2, 1), ('f_sparse_comp', 3, 2), ('a_max', 4, 2),
                               ('a_max', 5, 3), ('f_sparse_last', 6, 5), ('f_sparse_last', 7, 5)],
                   concat_node=[4, 5, 6, 7], score_func='sf_DisMult')]
REL = 1e-5


from parity import grad_errors, rel_err as _err  # max|a-b| / max|b|: relative to the tensor's own scale, no absolute floor


@pytest.fixture(scope="module")
def c1():
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.synth import CONFIGS, synth_kg
    dev = torch.device("cuda:0")
    N, R, T, D = CONFIGS["c1_fb15k237"]
    trip = synth_kg(N, R, T, seed=0)
    g = MRGraph.from_triples(N, trip, R, device=dev)
    return types.SimpleNamespace(dev=dev, N=N, R=R, T=T, D=D, trip=trip, g=g, E=2 * T, M=2 * T + N)


def test_c1_graph_segments_are_permutations(c1):
    g = c1.g
    for seg, total, keys in ((g.csr, c1.E, g.dst), (g.csc, c1.M, g.src_final), (g.rel, c1.M, g.et_final)):
        idx = seg.idx[:total].long()
        assert torch.equal(torch.sort(idx).values, torch.arange(total, device=c1.dev))       # each row once
        k = keys.long()[idx]
        assert bool((k[1:] >= k[:-1]).all())                                                  # grouped by key
        same = k[1:] == k[:-1]
        assert bool((idx[1:][same] > idx[:-1][same]).all())                                   # ascending ids inside
        ptr = seg.ptr.long()
        assert int(ptr[0]) == 0 and int(ptr[-1]) == total
        assert torch.equal(ptr[1:] - ptr[:-1], torch.bincount(keys.long(), minlength=seg.nseg))
    deg = torch.bincount(g.dst.long(), minlength=c1.N)
    assert torch.equal(g.in_deg.long(), deg)


def test_c1_a_sum_conserves_column_sums(c1):
    from mr_gnas_b200 import operations_lp as ops
    torch.manual_seed(0)
    x = torch.randn(c1.M, c1.D, device=c1.dev)
    out = ops.a_sum_op({'feature_dim': c1.D, 'drop_aggr': 0.0}).to(c1.dev)(c1.g, x, None)
    assert _err(out.double().sum(0), x.double().sum(0)) <= REL
    # and every destination individually, against index_add_ in fp64
    ref = torch.zeros(c1.N, c1.D, dtype=torch.float64, device=c1.dev)
    ref.index_add_(0, c1.g.dst.long(), x[:c1.E].double())
    ref += x[c1.E:].double()
    assert _err(out, ref) <= REL


def test_c1_a_max_matches_same_size_torch(c1):
    from mr_gnas_b200 import operations_lp as ops
    from mr_gnas_b200.functional import decode_arg
    g, E, N, D, dev = c1.g, c1.E, c1.N, c1.D, c1.dev
    torch.manual_seed(1)
    op = ops.a_max_op({'feature_dim': D}).to(dev)
    nn.init.xavier_normal_(op.linear.weight)
    op.linear.bias.data.normal_(0, 0.1)
    x = torch.relu(torch.randn(c1.M, D, device=dev)).requires_grad_(True)
    out = op(g, x, None)
    arg = decode_arg(g.last_arg).long()
    with torch.no_grad():
        msg = torch.relu(torch.addmm(op.linear.bias, x[:E], op.linear.weight.t()))          # [E, D] fp32 (checker)
        dst = g.dst.long()
        ref = torch.zeros(N, D, device=dev).scatter_reduce(0, dst.view(-1, 1).expand(E, D), msg, "amax",
                                                          include_self=False)
        deg = torch.bincount(dst, minlength=N)
        ref[deg == 0] = 0
        ref = ref + x[E:]
    assert _err(out, ref) <= REL
    assert bool((arg[deg == 0] == -1).all()) and bool((arg[deg > 0] >= 0).all())
    has = arg >= 0
    rows = torch.arange(N, device=dev).view(-1, 1).expand(N, D)
    assert bool((dst[arg[has]] == rows[has]).all())                       # the arg is an in-edge of its destination
    cols = torch.arange(D, device=dev).view(1, -1).expand(N, D)
    picked = msg[arg[has], cols[has]]
    best = (ref - x[E:].detach())[has]
    assert float((picked - best).abs().max()) <= 1e-5 * max(1.0, float(best.abs().max()))   # ... that attains the max
    # backward: gradient routed through the reported arg, exactly
    cot = torch.randn(N, D, device=dev)
    out.backward(cot)
    pos = (ref - x[E:].detach()) > 0
    gm = torch.zeros(E, D, device=dev, dtype=torch.float64)
    sel = has & pos
    gm.index_put_((arg[sel], cols[sel]), cot.double()[sel], accumulate=True)
    dx_ref = gm @ op.linear.weight.double()
    assert _err(x.grad[:E], dx_ref) <= REL
    assert _err(x.grad[E:], cot) <= REL
    assert _err(op.linear.weight.grad, gm.t() @ x[:E].detach().double()) <= REL
    assert _err(op.linear.bias.grad, gm.sum(0)) <= REL


def test_c1_gather_backward_matches_index_add(c1):
    from mr_gnas_b200 import functional as K
    g, D, dev = c1.g, c1.D, c1.dev
    torch.manual_seed(2)
    ent = torch.randn(c1.N, D, device=dev, requires_grad=True)
    rel = torch.randn(2 * c1.R + 1, D, device=dev, requires_grad=True)
    y, _ = K.GatherCompose.apply(ent, rel, g, 0)
    ref = ent.detach()[g.src_final.long()] - rel.detach()[g.et_final.long()]
    assert torch.equal(y, ref)                                             # one fp32 subtraction per element
    cot = torch.randn(c1.M, D, device=dev)
    y.backward(cot)
    de = torch.zeros(c1.N, D, dtype=torch.float64, device=dev).index_add_(0, g.src_final.long(), cot.double())
    dr = torch.zeros(2 * c1.R + 1, D, dtype=torch.float64, device=dev).index_add_(0, g.et_final.long(), -cot.double())
    assert _err(ent.grad, de) <= REL
    assert _err(rel.grad, dr) <= REL


def _model(c1, seed=0):
    from mr_gnas_b200.model_lp import Network
    from mr_gnas_b200.utils import weights_init
    args = types.SimpleNamespace(feature_dim=c1.D, drop_aggr=0.0, drop_op=0.0, gamma=40, embed_dim=c1.D,
                                 conve_hid_drop=0.0, feat_drop=0.0, num_filt=4, ker_sz=3, k_w=4, k_h=c1.D // 4)
    torch.manual_seed(seed)
    m = Network(c1.dev, README, c1.N, c1.R, c1.D, c1.D, 2 * c1.R + 1, nn.BCELoss(), 0.0, args)
    m.apply(weights_init)
    return m.to(c1.dev).train()


def _batch(c1, B=64):
    rng = np.random.RandomState(7)
    subj = torch.from_numpy(rng.randint(0, c1.N, B)).to(c1.dev)
    rel = torch.from_numpy(rng.randint(0, 2 * c1.R, B)).to(c1.dev)
    label = (torch.from_numpy(rng.rand(B, c1.N)) < 0.002).float().to(c1.dev) * 0.9 + 1.0 / c1.N
    return subj, rel, label


def test_c1_fused_cell_equals_modular_path_and_is_deterministic(c1):
    from mr_gnas_b200 import model_lp
    subj, rel, label = _batch(c1)
    runs = {}
    for tag, fused in (("fused", True), ("fused2", True), ("modular", False)):
        model_lp.USE_FUSED_CELL = fused
        m = _model(c1)
        loss = m._loss(c1.g, subj, rel, label)
        loss.backward()
        runs[tag] = (loss.detach().clone(), {k: p.grad.detach().clone() for k, p in m.named_parameters()
                                             if p.grad is not None})
        del m
    model_lp.USE_FUSED_CELL = True
    assert torch.equal(runs["fused"][0], runs["fused2"][0])
    for k, gk in runs["fused"][1].items():
        assert torch.equal(gk, runs["fused2"][1][k]), f"{k}: two runs differ (non-deterministic reduction?)"
    assert _err(runs["fused"][0], runs["modular"][0]) <= REL
    worst = max(grad_errors(runs["fused"][1], runs["modular"][1]))
    assert worst[0] <= 2e-5, worst


def test_c1_loss_falls_over_adam_steps(c1):
    subj, rel, label = _batch(c1)
    m = _model(c1)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        loss = m._loss(c1.g, subj, rel, label)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
