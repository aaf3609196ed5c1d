"""GPU parity for the remaining hot-path rows: DARTS MixedOp / LP supernet (cell_lp.py, model_search_lp.py),
NC operators / derived net / supernet (operations.py, model.py, cell.py, model_search.py) and CompGraphConv
(compgcn.py) -- all against golden vectors produced by the real reference code."""
import os
import types
from collections import namedtuple

import numpy as np
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
REL = 1e-5
Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func", defaults=(None,))


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


from parity import check_grads, grad_errors, rel_err as _err  # max|a-b| / max|b|: relative to the tensor's own scale, no absolute floor


def _check(name, a, b, tol=REL):
    assert tuple(a.shape) == tuple(b.shape), f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    e = _err(a, b)
    assert e <= tol, f"{name}: rel err {e:.3e} > {tol:.1e}"


DEV = "cuda:0"


def _lp_graph(gd):
    from mr_gnas_b200.graph import MRGraph
    return MRGraph.from_triples(gd["num_ent"], gd["triples"].numpy(), gd["num_rels"], device=DEV)


# ------------------------------------------------------------------------------ MixedOp (K9)
@pytest.mark.parametrize("tag", ["pre", "first", "middle", "last"])
def test_mixed_op_golden(golden_dir, tag):
    from mr_gnas_b200.cell_lp import MixedOp
    G = _load(golden_dir, "mixed_op.pt")
    g = _lp_graph(G["graph"])
    c = G["cases"][tag]
    mo = MixedOp(G["D"], 0.0, c["names"]).to(DEV)
    mo.load_state_dict(c["state"])
    mo.train()
    alpha = c["alpha"].to(DEV).requires_grad_(True)
    x = c["x"].to(DEV).requires_grad_(True)
    xin = c["xin"].to(DEV).requires_grad_(True)
    out = mo(torch.softmax(alpha, 0), g, x, xin)
    _check("out", out, c["out"])
    out.backward(c["cot"].to(DEV))
    _check("dalpha", alpha.grad, c["dalpha"])
    _check("dx", x.grad, c["dx"])
    if c["dxin"] is not None:
        _check("dxin", xin.grad if xin.grad is not None else torch.zeros_like(xin), c["dxin"])
    for k, p in mo.named_parameters():
        if c["dparams"][k] is not None:
            _check("d" + k, p.grad, c["dparams"][k])
    # running statistics of every candidate's BN were updated exactly once
    for k, v in mo.state_dict().items():
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(c["state"][k]) + 1


@pytest.mark.parametrize("names", [["pre_mult", "pre_sub", "pre_add"], ["pre_sub", "pre_mult"], ["pre_add"]])
@pytest.mark.parametrize("train", [True, False])
def test_mixed_pre_fused_equals_per_candidate_form(names, train):
    """a-11: the PRE-list MixedOp with ONE shared read of its inputs (mrg_mixed_pre_*: no candidate output is ever
    written) against the per-candidate form (K compositions + K BatchNorms + mixed sum) on the same module:
    output, both input gradients, dalpha, every BatchNorm gradient and the running statistics, at D = 200 with a
    ragged row count; and against an fp64 evaluation of the reference formula (cell_lp.py:25-33)."""
    from mr_gnas_b200 import functional as K
    from mr_gnas_b200.cell_lp import MixedOp
    torch.manual_seed(3)
    rows, D = 5003, 200
    x0, r0 = torch.randn(rows, D, device=DEV), torch.randn(rows, D, device=DEV)
    cot = torch.randn(rows, D, device=DEV)
    alpha0 = torch.randn(len(names), device=DEV)
    res = {}
    base = MixedOp(D, 0.0, names).to(DEV)
    for stack in base._ops:
        stack[-2].weight.data.uniform_(0.5, 1.5)
        stack[-2].bias.data.normal_(0, 0.3)
        stack[-2].running_mean.normal_(0, 0.5)
        stack[-2].running_var.uniform_(0.5, 2.0)
    state = {k: v.clone() for k, v in base.state_dict().items()}
    for fused in (True, False):
        K.MIXED_PRE_FUSED = fused
        mo = MixedOp(D, 0.0, names).to(DEV)
        mo.load_state_dict(state)
        mo.train(train)
        x, r, alpha = x0.clone().requires_grad_(True), r0.clone().requires_grad_(True), alpha0.clone().requires_grad_(True)
        out = mo(torch.softmax(alpha, 0), None, x, r)
        out.backward(cot)
        res[fused] = dict(out=out.detach(), dx=x.grad, dr=r.grad, dalpha=alpha.grad,
                          **{"d" + k: p.grad for k, p in mo.named_parameters()},
                          **{k: v.clone() for k, v in mo.state_dict().items() if "running" in k or "tracked" in k})
    K.MIXED_PRE_FUSED = True
    for k in res[True]:
        a, b = res[True][k], res[False][k]
        if a.dtype.is_floating_point:
            _check(k, a, b)
        else:
            assert torch.equal(a, b), k
    # fp64 statement of the reference formula
    x, r, alpha = x0.double().requires_grad_(True), r0.double().requires_grad_(True), alpha0.double().requires_grad_(True)
    w = torch.softmax(alpha, 0)
    tot = 0
    for k, name in enumerate(names):
        v = {"pre_mult": x * r, "pre_sub": x - r, "pre_add": x + r}[name]
        gam, bet = state[f"_ops.{k}.1.weight"].double(), state[f"_ops.{k}.1.bias"].double()
        if train:
            mu, var = v.mean(0), v.var(0, unbiased=False)
        else:
            mu, var = state[f"_ops.{k}.1.running_mean"].double(), state[f"_ops.{k}.1.running_var"].double()
        tot = tot + w[k] * torch.relu((v - mu) / torch.sqrt(var + 1e-5) * gam + bet)
    tot.backward(cot.double())
    _check("out vs fp64", res[True]["out"], tot.detach().float())
    _check("dx vs fp64", res[True]["dx"], x.grad.float())
    _check("dr vs fp64", res[True]["dr"], r.grad.float())
    _check("dalpha vs fp64", res[True]["dalpha"], alpha.grad.float())


def test_gather_rows_and_gather_few_match_torch_indexing():
    """The search network's gathers (model_search_lp.py:139-145,171): K.gather_rows (backward = the graph's segmented sum)
    and K.gather_few (backward = one-hot reduction GEMM) against torch indexing and its index_put backward in fp64."""
    from mr_gnas_b200 import functional as K
    from mr_gnas_b200.graph import MRGraph
    import oracle.mrg_oracle as O
    N, R, D = 700, 5, 200
    g = MRGraph.from_triples(N, O.synth_kg(N, R, 3000, seed=2), R, device=DEV)
    torch.manual_seed(4)
    ent = torch.randn(N, D, device=DEV, requires_grad=True)
    rel = torch.randn(2 * R + 1, D, device=DEV, requires_grad=True)
    cot = torch.randn(g.M, D, device=DEV)
    ye, yr = K.gather_rows(ent, g.src_final, g.csc), K.gather_rows(rel, g.et_final, g.rel)
    assert torch.equal(ye, ent[g.src_final.long()]) and torch.equal(yr, rel[g.et_final.long()])
    (ye * cot + yr * cot.flip(0)).sum().backward()
    e64, r64 = ent.detach().double().requires_grad_(True), rel.detach().double().requires_grad_(True)
    (e64[g.src_final.long()] * cot.double() + r64[g.et_final.long()] * cot.double().flip(0)).sum().backward()
    _check("d ent", ent.grad, e64.grad.float())
    _check("d rel", rel.grad, r64.grad.float())
    idx = torch.randint(0, 2 * R + 1, (33000,), device=DEV)
    cot2 = torch.randn(33000, D, device=DEV)
    rel.grad = None
    yf = K.gather_few(rel, idx)
    assert torch.equal(yf, rel[idx])
    (yf * cot2).sum().backward()
    r64.grad = None
    (r64[idx] * cot2.double()).sum().backward()
    _check("d rel (few rows, many gathers)", rel.grad, r64.grad.float())


# ------------------------------------------------------------------------------ LP supernet
def test_search_lp_golden(golden_dir):
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.model_search_lp import Network
    from mr_gnas_b200.utils import weights_init
    G = _load(golden_dir, "search_lp.pt")
    N, R, D, D0 = G["num_ent"], G["num_rels"], G["D"], G["D0"]
    # seeded construction reproduces the reference init (parameters AND alphas)
    torch.manual_seed(5)
    np.random.seed(5)
    from oracle.mrg_oracle import synth_kg
    synth_kg(N, R, 60, seed=13)  # (numpy RandomState is local; kept for symmetry with the generator)
    model = Network('cpu', N, R, 2, 1, 2, 2, D, D0, 2 * R + 1, 40, 0.0, 0.0)
    model.apply(weights_init)
    assert list(model.state_dict().keys()) == G["state_keys"]
    for k, v in model.state_dict().items():
        assert torch.equal(v, G["state0"][k]), k
    for a, b in zip(model.arch_parameters(), G["alphas0"]):
        assert torch.equal(a.detach(), b)
    assert str(model.show_genotypes()) == G["genotypes"]
    # forward / loss / grads on the device
    model = model.to(DEV)
    model._device = DEV
    alphas = [a.detach().to(DEV).requires_grad_(True) for a in G["alphas0"]]
    (model.alphas_zero_cell, model.alphas_first_cell, model.alphas_middle_cell, model.alphas_last_cell,
     model.alphas_final_cell) = alphas
    model._arch_parameters = alphas
    model.train()
    g = MRGraph.from_edges(G["src"], G["dst"], G["etype"], N, 2 * R + 1, device=DEV)
    assert np.array_equal(g.edge_norm.cpu().numpy(), G["norm"].view(-1).numpy())
    loss = model._loss(g, G["node_id"].to(DEV), G["src"].to(DEV), G["etype"].to(DEV), G["samples"].to(DEV),
                       G["labels"].to(DEV))
    _check("loss", loss.view(1), G["loss"].view(1))
    loss.backward()
    check_grads("search_lp", {k: p.grad for k, p in model.named_parameters()}, G["grads"], G["grads64"])
    check_grads("search_lp dalpha", {i: a.grad for i, a in enumerate(alphas)},
                {i: r for i, r in enumerate(G["dalphas"])}, {i: r for i, r in enumerate(G["dalphas64"])})


# ------------------------------------------------------------------------------ NC operators
@pytest.mark.parametrize("name,tc", [(n, True) for n in ['a_max', 'a_mean', 'a_sum', 'a_std', 'f_dense', 'f_sparse',
                                                         'f_dense_last', 'f_sparse_last']] + [('a_max', False)])
def test_nc_op_golden(golden_dir, name, tc):
    from mr_gnas_b200 import operations as ops
    from mr_gnas_b200 import operations_lp
    from mr_gnas_b200.graph import MRGraph
    G = _load(golden_dir, "ops_nc.pt")
    c = G["cases"][name]
    g = MRGraph.from_block(G["dst"], G["n_dst"], device=DEV)
    op = ops.MIXED_OPS[name]({'feature_dim': G["D"]}).to(DEV)
    op.load_state_dict(c["state"])
    x = c["x"].to(DEV).requires_grad_(True)
    xin = c["xin"].to(DEV).requires_grad_(True)
    operations_lp.USE_TENSOR_CORES = tc
    try:
        out = op(g, x, xin)
        _check("out", out, c["out"])
        out.backward(c["cot"].to(DEV))
    finally:
        operations_lp.USE_TENSOR_CORES = True
    _check("dx", x.grad, c["dx"])
    if c["dxin"] is not None:
        _check("dxin", xin.grad if xin.grad is not None else torch.zeros_like(xin), c["dxin"])
    for k, p in op.named_parameters():
        if c["dparams"][k] is not None:
            _check("d" + k, p.grad, c["dparams"][k])


# ------------------------------------------------------------------------------ NC networks
def _nc_blocks(G):
    from mr_gnas_b200.graph import MRBlock
    return [MRBlock.build(b["eids"], G["etype"][b["eids"]], b["local"], b["dst_nid"], device=DEV) for b in G["blocks"]]


def test_block_sampler_matches_golden_blocks(golden_dir):
    from mr_gnas_b200.graph import full_neighbor_blocks
    G = _load(golden_dir, "network_nc.pt")
    blocks = full_neighbor_blocks(G["src"].numpy(), G["dst"].numpy(), G["etype"].numpy(), G["seeds"].numpy(), 2,
                                  device=DEV)
    for b, ref in zip(blocks, G["blocks"]):
        assert torch.equal(b.edata['_ID'].cpu(), ref["eids"])
        assert torch.equal(b.dst.cpu().long(), ref["local"])
        assert torch.equal(b.dstdata['_ID'].cpu(), ref["dst_nid"])


@pytest.mark.parametrize("op_norm", [True, False])
def test_network_nc_golden(golden_dir, op_norm):
    from mr_gnas_b200.model import Network
    G = _load(golden_dir, "network_nc.pt")
    ref = G["derived_norm%d" % int(op_norm)]
    args = types.SimpleNamespace(feature_dim=G["D"], op_norm=op_norm)
    model = Network(DEV, eval(G["genotype"]), G["N"], G["C"], G["ET"], 2, 1, 2, G["D"], G["D0"], G["NB"],
                    nn.CrossEntropyLoss(), args)
    assert list(model.state_dict().keys()) == ref["state_keys"]
    model.load_state_dict(ref["state0"])
    model = model.to(DEV).train()
    logits = model(G["trip_index"].to(DEV), _nc_blocks(G))
    _check("logits", logits, ref["logits"])
    loss = nn.CrossEntropyLoss()(logits, G["labels"].to(DEV))
    _check("loss", loss.view(1), ref["loss"].view(1))
    loss.backward()
    check_grads("network_nc", {k: p.grad for k, p in model.named_parameters()}, ref["grads"], ref["grads64"])


def test_search_nc_golden(golden_dir):
    from mr_gnas_b200.model_search import Network
    from mr_gnas_b200.utils import weights_init
    G = _load(golden_dir, "network_nc.pt")
    ref = G["search"]
    torch.manual_seed(9)
    model = Network('cpu', G["N"], G["C"], G["ET"], 2, 1, 2, G["D"], G["D0"], G["NB"], 0.0)
    model.apply(weights_init)
    assert list(model.state_dict().keys()) == ref["state_keys"]
    for k, v in model.state_dict().items():
        assert torch.equal(v, ref["state0"][k]), k
    for a, b in zip(model.arch_parameters(), ref["alphas0"]):
        assert torch.equal(a.detach(), b)
    assert str(model.show_genotypes()) == ref["genotypes"]
    model = model.to(DEV)
    alphas = [a.detach().to(DEV).requires_grad_(True) for a in ref["alphas0"]]
    model.alphas_zero_cell, model.alphas_first_cell, model.alphas_middle_cell, model.alphas_last_cell = alphas
    model._arch_parameters = alphas
    model.train()
    logits = model(G["trip_index"].to(DEV), _nc_blocks(G))
    _check("logits", logits, ref["logits"])
    loss = nn.CrossEntropyLoss()(logits, G["labels"].to(DEV))
    loss.backward()
    check_grads("search_nc", {k: p.grad for k, p in model.named_parameters()}, ref["grads"], ref["grads64"])
    check_grads("search_nc dalpha", {i: a.grad for i, a in enumerate(alphas)},
                {i: r for i, r in enumerate(ref["dalphas"])}, {i: r for i, r in enumerate(ref["dalphas64"])})


# ------------------------------------------------------------------------------ CompGraphConv (K10)
@pytest.mark.parametrize("comp", ["sub", "mul", "ccorr"])
def test_compgcn_golden(golden_dir, comp):
    from mr_gnas_b200.compgcn import CompGraphConv
    G = _load(golden_dir, "compgcn.pt")
    gd, c = G["graph"], G["cases"][comp]
    g = _lp_graph(gd)
    E = g.E
    g.edata['etype'] = g.edata['e_type']
    g.edata['in_edges_mask'] = torch.arange(E, device=DEV) < E // 2
    g.edata['out_edges_mask'] = torch.arange(E, device=DEV) >= E // 2
    layer = CompGraphConv(G["Din"], G["Dout"], comp_fn=comp, batchnorm=True, dropout=0.0).to(DEV)
    layer.load_state_dict(c["state"])
    layer.train()
    h = c["h"].to(DEV).requires_grad_(True)
    r = c["r"].to(DEV).requires_grad_(True)
    n_out, r_out = layer(g, h, r)
    _check("n_out", n_out, c["n_out"])
    _check("r_out", r_out, c["r_out"])
    ((n_out * c["c1"].to(DEV)).sum() + (r_out * c["c2"].to(DEV)).sum()).backward()
    # gradients against the fp64 run of the same reference layer (bar: max(1e-5, 4 x the fp32 reference's own
    # error); exactly-zero true gradients -- loop_rel and W_S.bias under `sub`, any bias feeding the BatchNorm --
    # must be negligible): tests/parity.py::check_grads
    T = c["truth64"]
    ours = {"dh": h.grad, "dr": r.grad, **{"d" + k: p.grad for k, p in layer.named_parameters()}}
    ref32 = {"dh": c["dh"], "dr": c["dr"], **{"d" + k: v for k, v in c["dparams"].items()}}
    truth = {"dh": T["dh"], "dr": T["dr"], **{"d" + k: v for k, v in T["dparams"].items()}}
    check_grads("compgcn " + comp, ours, ref32, truth)


# ------------------------------------------------------------------------------ search-script graph pipeline (8f rank 3)
@pytest.mark.parametrize("N,R,T,S", [(500, 5, 3000, 400), (40943, 11, 86835, 30000)], ids=["small", "c3"])
def test_device_sampler_reproduces_reference_pipeline(N, R, T, S):
    """utils_rgcn.sample_search_graph (device: torch.unique relabel, vectorised negative sampling, packed-key radix
    sort by (rel, dst, src), K0 graph build) against the oracle restatement of generate_sampled_graph_and_labels --
    itself pinned draw-for-draw to the REAL utils/utils_rgcn.py:79-204 (tests/golden/config_c3.pt checksums) -- on
    the SAME random draws: every array must be bit-identical (integer graph arrays, fp32 norms, samples, labels)."""
    from mr_gnas_b200.utils_rgcn import sample_search_graph
    from oracle.mrg_oracle import sample_search_graph as oracle_sampler, synth_kg
    trip = synth_kg(N, R, T, seed=0)
    neg_rate, split_size = 10, 0.5
    np.random.seed(0)
    ref = oracle_sampler(trip, S, split_size, R, neg_rate)
    # the same draws, taken in the reference's order from the same numpy stream
    np.random.seed(0)
    edges = np.random.choice(np.arange(T), S, replace=False)
    n = len(np.unique((trip[edges, 0], trip[edges, 2])))
    values = np.random.randint(n, size=S * neg_rate)
    choices = np.random.uniform(size=S * neg_rate)
    keep = np.random.choice(np.arange(S), size=int(S * split_size), replace=False)
    d = sample_search_graph(trip, S, split_size, R, neg_rate, device=DEV,
                            draws={"edges": edges, "values": values, "choices": choices, "keep": keep})
    for key in ("uniq_v", "src", "dst", "etype", "samples"):
        assert np.array_equal(d[key].cpu().numpy(), ref[key]), key
    assert np.array_equal(d["labels"].cpu().numpy(), ref["labels"])
    assert np.array_equal(d["node_norm"].cpu().numpy(), ref["node_norm"])
    assert np.array_equal(d["norm"].cpu().numpy(), ref["norm"])
    g = d["g"]
    assert g.N == ref["num_nodes"] and g.E == len(ref["src"]) and tuple(g.edata['norm'].shape) == (g.E, 1)
    # its own generator: structurally valid sample (sizes, label counts, relabelling is a bijection onto [0, n))
    gen = torch.Generator(device=DEV).manual_seed(3)
    d2 = sample_search_graph(trip, S, split_size, R, neg_rate, device=DEV, generator=gen)
    assert d2["samples"].shape == (S * (neg_rate + 1), 3) and int(d2["labels"].sum()) == S
    assert d2["g"].E == 2 * int(S * split_size) and int(d2["src"].max()) < d2["g"].N
    uv = d2["uniq_v"].cpu().numpy()
    assert np.all(np.diff(uv) > 0) and uv.max() < N
    key = (d2["etype"] * d2["g"].N + d2["dst"]) * d2["g"].N + d2["src"]
    assert bool((key[1:] >= key[:-1]).all())          # ordered by (rel, dst, src)


def test_compgcn_sub_node_level_equals_edge_level():
    """The node-level reorder of CompGraphConv(sub) (two node GEMMs + one gather-sum, no [E, D] tensor) against the
    edge-level evaluation of the same layer (composition kernel, masked edge-tile GEMMs, segmented sum) on a
    3,000-node / 40,000-edge graph: outputs and every gradient; reference: models/compgcn.py:62-100."""
    from mr_gnas_b200 import compgcn
    from mr_gnas_b200.graph import MRGraph
    from oracle.mrg_oracle import synth_kg
    from parity import check_grads_pair
    N, R, T, Din, Dout = 3000, 7, 20000, 64, 128
    trip = synth_kg(N, R, T, seed=4)
    g = MRGraph.from_triples(N, trip, R, device=DEV)
    E = g.E
    g.edata['etype'] = g.edata['e_type']
    g.edata['in_edges_mask'] = torch.arange(E, device=DEV) < E // 2
    g.edata['out_edges_mask'] = torch.arange(E, device=DEV) >= E // 2
    torch.manual_seed(0)
    layer = compgcn.CompGraphConv(Din, Dout, comp_fn='sub', batchnorm=True, dropout=0.0).to(DEV).train()
    h0, r0 = torch.randn(N, Din, device=DEV), torch.randn(2 * R, Din, device=DEV)
    c1, c2 = torch.randn(N, Dout, device=DEV), torch.randn(2 * R, Dout, device=DEV)
    res = {}
    for node_level in (True, False):
        compgcn.NODE_LEVEL_SUB = node_level
        layer.zero_grad()
        h, r = h0.clone().requires_grad_(True), r0.clone().requires_grad_(True)
        n_out, r_out = layer(g, h, r)
        ((n_out * c1).sum() + (r_out * c2).sum()).backward()
        res[node_level] = (n_out.detach(), {"dh": h.grad, "dr": r.grad,
                                            **{k: p.grad.clone() for k, p in layer.named_parameters()}})
    compgcn.NODE_LEVEL_SUB = True
    assert _err(res[True][0], res[False][0]) <= 1e-5
    check_grads_pair("compgcn sub node-level vs edge-level", res[True][1], res[False][1])
