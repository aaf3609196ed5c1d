"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the REAL reference.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):    python oracle/make_golden.py

The reference's Python (models/operations_lp.py, operations.py, model_lp.py, cell_lp.py,
model_search_lp.py, compgcn.py, utils/process_data.py, utils/data_set.py, utils.weights_init)
is imported unmodified; ``oracle/dgl_stub.py`` stands in for DGL.  Each fixture stores the
seeded inputs, the reference ``state_dict``, outputs and gradients, so the tests can replay
it through (a) the oracle restatement on CPU and (b) the CUDA path on the GPU box.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MRG_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import dgl_stub  # noqa: E402
from oracle.mrg_oracle import synth_kg  # noqa: E402

dgl_stub.install()
sys.path.insert(0, REF)
import models.operations_lp as ref_lp  # noqa: E402
import models.operations as ref_nc  # noqa: E402
import models.model_lp as ref_model_lp  # noqa: E402
import models.cell_lp as ref_cell_lp  # noqa: E402
import models.model_search_lp as ref_search_lp  # noqa: E402
import models.compgcn as ref_compgcn  # noqa: E402
from configs.genotypes import Genotype  # noqa: E402
from utils.utils import weights_init  # noqa: E402
from utils.process_data import process  # noqa: E402
from utils.data_set import TrainDataset  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
README_GENOTYPE = ("[Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_sparse_comp', 2, 1), ('f_sparse_comp', 3, 2), "
                   "('a_max', 4, 2), ('a_max', 5, 3), ('f_sparse_last', 6, 5), ('f_sparse_last', 7, 5)], "
                   "concat_node=[4, 5, 6, 7], score_func='sf_DisMult')]")


def ref_build_graph(num_ent, data, num_rels):
    """Verbatim call sequence of train/mr_lp_train.py:77-89 against the stub graph."""
    import dgl
    g = dgl.DGLGraph()
    g.add_nodes(num_ent)
    g.add_edges(data[:, 0], data[:, 2])
    g.add_edges(data[:, 2], data[:, 0])
    in_deg = g.in_degrees(range(g.number_of_nodes())).cpu().float().numpy()
    norm = in_deg ** -0.5
    norm[np.isinf(norm)] = 0
    g.ndata['n_norm'] = torch.tensor(norm)
    g.apply_edges(lambda edges: {'norm': edges.dst['n_norm'] * edges.src['n_norm']})
    edge_type = torch.tensor(np.concatenate([data[:, 1], data[:, 1] + num_rels]))
    g.edata['e_type'] = edge_type
    return g


def _sd(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def _grads(module):
    return {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in module.named_parameters()}


def make_graph_case(N, R, T, seed):
    trip = synth_kg(N, R, T, seed=seed)
    trip[:3, 2] = trip[:3, 0]          # a few self-edges
    trip[3:6] = trip[0]                # multi-edges
    g = ref_build_graph(N, trip, R)
    src, dst, _ = g.edges(form='all')
    return trip, g, {"triples": torch.from_numpy(trip), "num_ent": N, "num_rels": R,
                     "src": src.clone(), "dst": dst.clone(), "etype": g.edata['e_type'].clone(),
                     "norm": g.edata['norm'].clone(), "in_deg": g.in_degrees().clone()}


def gen_ops_lp():
    torch.manual_seed(1)
    N, R, T, D = 37, 4, 90, 16
    trip, g, gd = make_graph_case(N, R, T, seed=3)
    E = g.num_edges()
    M = E + N
    cases = {}
    for name in ['pre_mult', 'pre_sub', 'pre_add', 'f_zero', 'f_identity', 'f_dense', 'f_dense_comp', 'f_comp',
                 'f_sparse', 'f_sparse_comp', 'f_dense_last', 'f_sparse_last', 'a_max', 'a_mean', 'a_sum']:
        op = ref_lp.MIXED_OPS[name]({'feature_dim': D, 'drop_aggr': 0.0})
        op.apply(weights_init)
        rows = N if name.endswith('_last') else M
        x = torch.randn(rows, D)
        if name.startswith('a_'):
            x = torch.relu(x)  # aggregator inputs are post-ReLU in the cell; gives exact-zero ties
            x[:E:7] = x[1:E:7][: x[:E:7].shape[0]]  # duplicate rows -> ties between edges
        xin = torch.randn(rows, D)
        x.requires_grad_(True)
        xin.requires_grad_(True)
        out = op(g, x, xin)
        cot = torch.randn_like(out)
        out.backward(cot)
        cases[name] = {"x": x.detach().clone(), "xin": xin.detach().clone(), "state": _sd(op), "out": out.detach().clone(),
                       "cot": cot, "dx": x.grad.clone() if x.grad is not None else None,
                       "dxin": xin.grad.clone() if xin.grad is not None else None, "dparams": _grads(op),
                       "arg": g.last_arg.clone() if name == 'a_max' else None}
    torch.save({"graph": gd, "D": D, "cases": cases}, os.path.join(OUT, "ops_lp.pt"))


def gen_ops_nc():
    torch.manual_seed(2)
    n_dst, Eb, D = 23, 80, 8
    rng = np.random.RandomState(5)
    dst = rng.randint(0, n_dst - 3, size=Eb)  # last 3 dst nodes isolated
    src = rng.randint(0, 50, size=Eb)
    g = dgl_stub.StubGraph(n_dst, src, dst)
    cases = {}
    for name in ['a_max', 'a_mean', 'a_sum', 'a_std', 'f_dense', 'f_sparse', 'f_dense_last', 'f_sparse_last']:
        op = ref_nc.MIXED_OPS[name]({'feature_dim': D})
        op.apply(weights_init)
        rows = Eb
        x = torch.randn(rows, D, requires_grad=True)
        xin = torch.randn(rows, D, requires_grad=True)
        out = op(g, x, xin)
        cot = torch.randn_like(out)
        out.backward(cot)
        cases[name] = {"x": x.detach().clone(), "xin": xin.detach().clone(), "state": _sd(op), "out": out.detach().clone(),
                       "cot": cot, "dx": x.grad.clone() if x.grad is not None else None,
                       "dxin": xin.grad.clone() if xin.grad is not None else None, "dparams": _grads(op)}
    torch.save({"dst": torch.from_numpy(dst), "src": torch.from_numpy(src), "n_dst": n_dst, "D": D, "cases": cases},
               os.path.join(OUT, "ops_nc.pt"))


def _lp_args(D):
    return types.SimpleNamespace(feature_dim=D, drop_aggr=0.0, drop_op=0.0, gamma=40, embed_dim=D,
                                 conve_hid_drop=0.0, feat_drop=0.0, num_filt=4, ker_sz=3, k_w=4, k_h=D // 4)


def gen_network_lp():
    N, R, T, D, D0, B = 61, 5, 160, 16, 12, 8
    trip, g, gd = make_graph_case(N, R, T, seed=7)
    torch.manual_seed(0)
    np.random.seed(0)
    genotype = eval(README_GENOTYPE)
    model = ref_model_lp.Network('cpu', genotype, N, R, D, D0, 2 * R + 1, nn.BCELoss(), 0.0, _lp_args(D))
    model.apply(weights_init)
    state0 = _sd(model)
    triplets = process({'train': trip, 'valid': trip[:5], 'test': trip[:5]}, R)
    ds = TrainDataset(triplets['train'], N, types.SimpleNamespace(lbl_smooth=0.1))
    items = [ds[i] for i in range(B)]
    tr = torch.stack([it[0] for it in items])
    labels = torch.stack([it[1] for it in items])
    subj, rel = tr[:, 0], tr[:, 1]
    model.train()
    pred = model(g, subj, rel)
    loss = model.criterion(pred, labels)
    loss.backward()
    state1 = _sd(model)  # running stats after one training forward
    # a few SGD-free Adam steps to pin a short loss curve
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = [loss.item()]
    grads = _grads(model)
    opt.step()
    for _ in range(3):
        opt.zero_grad()
        l = model.criterion(model(g, subj, rel), labels)
        l.backward()
        opt.step()
        losses.append(l.item())
    model.eval()
    with torch.no_grad():
        pred_eval = model(g, subj, rel)
    torch.save({"graph": gd, "genotype": README_GENOTYPE, "dims": {"N": N, "R": R, "D": D, "D0": D0, "B": B},
                "state0": state0, "state1": state1, "subj": subj, "rel": rel, "labels": labels,
                "pred": pred.detach(), "loss": loss.detach(), "grads": grads, "losses": losses,
                "state_eval": _sd(model), "pred_eval": pred_eval,
                "train_items": [(list(it['triple']), list(it['label'])) for it in triplets['train'][:B]],
                "state_keys": list(state0.keys())},
               os.path.join(OUT, "network_lp.pt"))


def gen_mixed_op():
    torch.manual_seed(4)
    N, R, T, D = 29, 3, 70, 8
    trip, g, gd = make_graph_case(N, R, T, seed=11)
    E = g.num_edges()
    M = E + N
    cases = {}
    for tag, names, rows in [("pre", ref_lp.PRE_OPS, M), ("first", ref_lp.FIRST_OPS, M),
                             ("middle", ref_lp.MIDDLE_OPS, M), ("last", ref_lp.LAST_OPS, N)]:
        mo = ref_cell_lp.MixedOp(D, 0.0, names)
        mo.apply(weights_init)
        mo.train()
        alpha = (1e-1 * torch.randn(len(names))).requires_grad_(True)
        w = torch.softmax(alpha, 0)
        x = torch.randn(rows, D)
        if tag == "middle":
            x = torch.relu(x)
        x.requires_grad_(True)
        xin = torch.randn(rows, D, requires_grad=True)
        out = mo(w, g, x, xin)
        cot = torch.randn_like(out)
        out.backward(cot)
        cases[tag] = {"names": list(names), "alpha": alpha.detach().clone(), "x": x.detach().clone(),
                      "xin": xin.detach().clone(), "state": _sd(mo), "out": out.detach().clone(), "cot": cot,
                      "dalpha": alpha.grad.clone(), "dx": x.grad.clone(),
                      "dxin": xin.grad.clone() if xin.grad is not None else None, "dparams": _grads(mo)}
    torch.save({"graph": gd, "D": D, "cases": cases}, os.path.join(OUT, "mixed_op.pt"))


def gen_search_lp():
    """Supernet forward/loss (model_search_lp.py:131-194) on a small 'sampled' graph."""
    N, R, T, D, D0 = 41, 3, 60, 8, 6
    torch.manual_seed(5)
    np.random.seed(5)
    trip = synth_kg(N, R, T, seed=13)
    # search graph: edges sorted by (rel, dst, src) -- utils/utils_rgcn.py:139-152
    s, r, o = trip[:, 0], trip[:, 1], trip[:, 2]
    src = np.concatenate((s, o)); dst = np.concatenate((o, s)); rel = np.concatenate((r, r + R))
    edges = sorted(zip(rel, dst, src))
    rel, dst, src = np.array(edges).transpose()
    g = dgl_stub.StubGraph(N, src, dst)
    in_deg = g.in_degrees().float().numpy()
    nn_ = in_deg ** -0.5
    nn_[np.isinf(nn_)] = 0
    node_norm = torch.from_numpy(nn_).view(-1, 1)
    g.ndata['norm'] = node_norm
    g.apply_edges(lambda edges: {'norm': edges.dst['norm'] * edges.src['norm']})  # mr_lp_search.py:30-36
    model = ref_search_lp.Network('cpu', N, R, 2, 1, 2, 2, D, D0, 2 * R + 1, 40, 0.0, 0.0)
    model.apply(weights_init)
    model.train()
    node_id = torch.arange(N).view(-1, 1)
    src_in = torch.from_numpy(src).long()
    edge_type = torch.from_numpy(rel).long()
    neg = trip.copy()
    neg[:, 2] = np.random.randint(0, N, size=T)
    samples = torch.from_numpy(np.concatenate([trip, neg])).long()
    labels = torch.cat([torch.ones(T), torch.zeros(T)])
    state0 = _sd(model)
    alphas0 = [a.detach().clone() for a in model.arch_parameters()]
    loss = model._loss(g, node_id, src_in, edge_type, samples, labels)
    loss.backward()
    g.edata['norm'] = g.edata['norm'].double()
    grads64, dalphas64, loss64 = _truth64(model, state0, lambda m: m._loss(g, node_id, src_in, edge_type, samples,
                                                                            labels.double()), alphas0)
    g.edata['norm'] = g.edata['norm'].float()
    torch.save({"grads64": grads64, "dalphas64": dalphas64, "loss64": loss64, "num_ent": N, "num_rels": R, "D": D, "D0": D0, "src": src_in, "dst": torch.from_numpy(dst).long(),
                "etype": edge_type, "norm": g.edata['norm'].clone(), "node_id": node_id, "samples": samples,
                "labels": labels, "state0": state0, "alphas0": alphas0, "loss": loss.detach(),
                "grads": _grads(model),
                "dalphas": [a.grad.clone() if a.grad is not None else None for a in model.arch_parameters()],
                "genotypes": str(model.show_genotypes()), "state_keys": list(state0.keys())},
               os.path.join(OUT, "search_lp.pt"))


def gen_compgcn():
    torch.manual_seed(6)
    N, R, T, Din, Dout = 31, 3, 64, 8, 12
    trip, g0, gd = make_graph_case(N, R, T, seed=17)
    E = g0.num_edges()
    g0.edata['etype'] = g0.edata['e_type'].long()
    g0.edata['in_edges_mask'] = torch.arange(E) < E // 2
    g0.edata['out_edges_mask'] = torch.arange(E) >= E // 2
    g0.edata['norm'] = g0.edata['norm'].view(-1)

    class _Scope:
        def __enter__(self): return self
        def __exit__(self, *a): return False
    g0.local_scope = lambda: _Scope()
    orig_apply = g0.apply_edges

    def apply_edges(fn):
        if isinstance(fn, dgl_stub._Desc):
            u, e = g0.ndata[fn.a][g0._src], g0.edata[fn.b]
            g0.edata['comp_h'] = u - e if fn.kind == 'u_sub_e' else u * e
        else:
            orig_apply(fn)
    g0.apply_edges = apply_edges
    # utils/utils.py:285-301 uses the removed torch.rfft; restate with torch.fft for the ccorr case only
    ref_compgcn.ccorr = lambda a, b: torch.fft.irfft(
        torch.conj(torch.fft.rfft(a, dim=-1)) * torch.fft.rfft(b, dim=-1), n=a.shape[-1], dim=-1)
    cases = {}
    for comp in ['sub', 'mul', 'ccorr']:
        layer = ref_compgcn.CompGraphConv(Din, Dout, comp_fn=comp, batchnorm=True, dropout=0.0)
        layer.apply(weights_init)
        layer.train()
        h = torch.randn(N, Din, requires_grad=True)
        r = torch.randn(2 * R, Din, requires_grad=True)
        n_out, r_out = layer(g0, h, r)
        c1, c2 = torch.randn_like(n_out), torch.randn_like(r_out)
        (n_out * c1).sum().add((r_out * c2).sum()).backward()
        import copy
        l64 = copy.deepcopy(layer).double().train()
        l64.zero_grad()
        h64, r64 = h.detach().double().requires_grad_(True), r.detach().double().requires_grad_(True)
        g0.edata['norm'] = g0.edata['norm'].double()
        torch.set_default_dtype(torch.float64)      # compgcn.py:79 allocates with the default dtype
        try:
            n64, ro64 = l64(g0, h64, r64)
            (n64 * c1.double()).sum().add((ro64 * c2.double()).sum()).backward()
        finally:
            torch.set_default_dtype(torch.float32)
        g0.edata['norm'] = g0.edata['norm'].float()
        truth = {"dh": h64.grad.clone(), "dr": r64.grad.clone(), "dparams": _grads(l64)}
        cases[comp] = {"truth64": truth,
                       "h": h.detach().clone(), "r": r.detach().clone(), "state": _sd(layer), "n_out": n_out.detach(),
                       "r_out": r_out.detach(), "c1": c1, "c2": c2, "dh": h.grad.clone(), "dr": r.grad.clone(),
                       "dparams": _grads(layer)}
    torch.save({"graph": gd, "Din": Din, "Dout": Dout, "cases": cases}, os.path.join(OUT, "compgcn.pt"))


NC_GENOTYPE = ("[Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_dense', 2, 1), ('f_sparse', 3, 2), ('f_identity', 4, 3), "
               "('a_sum', 5, 2), ('a_sum', 6, 3), ('a_mean', 7, 4), ('f_dense_last', 8, 7), ('f_sparse_last', 9, 7), "
               "('f_sparse_last', 10, 5)], concat_node=[5, 6, 7, 8, 9, 10]), Genotype(alpha_cell=[('pre_sub', 1, 0), "
               "('f_sparse', 2, 1), ('f_identity', 3, 2), ('f_identity', 4, 1), ('a_max', 5, 2), ('a_mean', 6, 3), "
               "('a_mean', 7, 4), ('f_sparse_last', 8, 7), ('f_sparse_last', 9, 8), ('f_identity', 10, 9)], "
               "concat_node=[5, 6, 7, 8, 9, 10])]")  # train/mr_nc_train.py:220 default


def _nc_blocks(src, dst, etype, seeds, layers):
    """Full-neighbour blocks as DGL's MultiLayerFullNeighborSampler(return_eids=True) yields them: block edges in
    ascending parent edge id, destinations = the layer's frontier, outermost block first."""
    blocks, frontier = [], np.asarray(seeds)
    for _ in range(layers):
        eids = np.nonzero(np.isin(dst, frontier))[0]
        pos = {int(n): i for i, n in enumerate(frontier)}
        local = np.array([pos[int(d)] for d in dst[eids]], dtype=np.int64)
        b = dgl_stub.StubGraph(len(frontier), src[eids], local)
        b.edata['_ID'] = torch.from_numpy(eids).long()
        b.edata['_TYPE'] = torch.from_numpy(etype[eids]).long()
        b.ndata['_ID'] = torch.from_numpy(frontier).long()
        blocks.append((b, eids, local, frontier.copy()))
        frontier = np.unique(np.concatenate([frontier, src[eids]]))
    return blocks[::-1]


def gen_network_nc():
    import models.model as ref_model_nc
    import models.model_search as ref_search_nc
    from collections import namedtuple
    G2 = namedtuple('Genotype', 'alpha_cell concat_node score_func', defaults=(None,))
    N, ET, E, D, D0, C, NB = 70, 6, 400, 16, 8, 3, 5
    rng = np.random.RandomState(23)
    src, dst = rng.randint(0, N, E), rng.randint(0, N - 5, E)
    etype = rng.randint(0, ET, E)
    seeds = np.sort(rng.choice(N - 5, 7, replace=False))
    blocks = _nc_blocks(src, dst, etype, seeds, 2)
    trip_index = torch.from_numpy(np.stack([np.arange(E), src, dst], 1)).long()
    labels = torch.from_numpy(rng.randint(0, C, len(seeds))).long()
    block_dump = [{"eids": torch.from_numpy(e), "local": torch.from_numpy(l), "dst_nid": torch.from_numpy(f)}
                  for (_, e, l, f) in blocks]
    out = {"N": N, "ET": ET, "D": D, "D0": D0, "C": C, "NB": NB, "src": torch.from_numpy(src), "dst": torch.from_numpy(dst),
           "etype": torch.from_numpy(etype), "seeds": torch.from_numpy(seeds), "blocks": block_dump,
           "trip_index": trip_index, "labels": labels, "genotype": NC_GENOTYPE}
    for op_norm in (True, False):
        torch.manual_seed(8)
        args = types.SimpleNamespace(feature_dim=D, op_norm=op_norm)
        geno = eval(NC_GENOTYPE, {"Genotype": G2})
        model = ref_model_nc.Network('cpu', geno, N, C, ET, 2, 1, 2, D, D0, NB, nn.CrossEntropyLoss(), args)
        model.apply(weights_init)
        model.train()
        state0 = _sd(model)
        logits = model(trip_index, [b for (b, _, _, _) in blocks])
        loss = nn.CrossEntropyLoss()(logits, labels)
        loss.backward()
        g64, _, l64 = _truth64(model, state0, lambda m: nn.CrossEntropyLoss()(
            m(trip_index, [b for (b, _, _, _) in blocks]), labels))
        out["derived_norm%d" % int(op_norm)] = {"state0": state0, "logits": logits.detach(), "loss": loss.detach(),
                                                 "grads": _grads(model), "state_keys": list(state0.keys()),
                                                 "grads64": g64, "loss64": l64}
    torch.manual_seed(9)
    sm = ref_search_nc.Network('cpu', N, C, ET, 2, 1, 2, D, D0, NB, 0.0)
    sm.apply(weights_init)
    sm.train()
    state0 = _sd(sm)
    alphas0 = [a.detach().clone() for a in sm.arch_parameters()]
    logits = sm(trip_index, [b for (b, _, _, _) in blocks])
    loss = nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    import configs.genotypes as cg
    g64, da64, l64 = _truth64(sm, state0, lambda m: nn.CrossEntropyLoss()(
        m(trip_index, [b for (b, _, _, _) in blocks]), labels), alphas0)
    out["search"] = {"grads64": g64, "dalphas64": da64, "loss64": l64,
                     "state0": state0, "alphas0": alphas0, "logits": logits.detach(), "loss": loss.detach(),
                     "grads": _grads(sm), "dalphas": [a.grad.clone() for a in sm.arch_parameters()],
                     "genotypes": str(sm.show_genotypes()), "state_keys": list(state0.keys())}
    torch.save(out, os.path.join(OUT, "network_nc.pt"))


# ------------------------------------------------------------------------------------------------------------
# BASELINE.json configuration shapes, run through the REAL reference; results stored as compact summaries
# (oracle/summary.py) so the fixtures stay small.  Inputs are regenerated from seeds by the tests (checksums stored).
# ------------------------------------------------------------------------------------------------------------
class _no_float_cast:
    """MixedOp casts every candidate output with .float() (cell_lp.py:30, cell.py:28): a no-op inside fp64 truth runs."""

    def __enter__(self):
        self._f = torch.Tensor.float
        torch.Tensor.float = lambda t, *a, **k: t

    def __exit__(self, *exc):
        torch.Tensor.float = self._f
        return False


def _truth64(model, state0, run, alphas0=None):
    """Gradients of the SAME reference modules evaluated in float64 from `state0`: run(model64) -> loss.
    -> (grads64, dalphas64 or None).  Consumes no random numbers."""
    import copy
    m = copy.deepcopy(model)
    m.load_state_dict(state0)
    m = m.double().train()
    m.zero_grad()
    if alphas0 is not None:
        for a, a0 in zip(m.arch_parameters(), alphas0):
            a.data.copy_(a0)
            a.grad = None
    with _no_float_cast():
        loss = run(m)
        loss.backward()
    dal = [a.grad.clone() if a.grad is not None else None for a in m.arch_parameters()] if alphas0 is not None else None
    return {k: (g.double() if g is not None else None) for k, g in _grads(m).items()}, dal, loss.detach().clone()


def _buffers(module):
    return {k: v.detach().clone() for k, v in module.state_dict().items() if "running" in k or "num_batches" in k}


def _summ_all(prefix, named):
    from oracle.summary import summarize
    return {k: (summarize(prefix + k, v) if v is not None else None) for k, v in named.items()}


def gen_config_c1():
    """configs[0]/C1: README genotype LP training step on the FB15k-237-shaped KG, full size, no rescaling of any
    parameter: N=14,541 R=237 T=272,115 D=200 B=256, label smoothing 0.1 (mr_lp_train.py:222-246)."""
    import time
    from oracle.summary import checksum, positions, summarize
    N, R, T, D, B = 14541, 237, 272115, 200, 256
    trip = synth_kg(N, R, T, seed=0)
    g = ref_build_graph(N, trip, R)
    torch.manual_seed(0)
    np.random.seed(0)
    model = ref_model_lp.Network('cpu', eval(README_GENOTYPE), N, R, D, D, 2 * R + 1, nn.BCELoss(), 0.0, _lp_args(D))
    model.apply(weights_init)
    state0 = _sd(model)
    items = process({'train': trip, 'valid': trip[:0], 'test': trip[:0]}, R)['train'][:B]
    ds = TrainDataset(items, N, types.SimpleNamespace(lbl_smooth=0.1))
    rows = [ds[i] for i in range(B)]
    tr, labels = torch.stack([r[0] for r in rows]), torch.stack([r[1] for r in rows])
    model.train()
    seen = {}
    hook = model.score_func.register_forward_hook(lambda mod, inp, out: seen.update(inp=[t.detach() for t in inp]))
    t0 = time.time()
    pred = model(g, tr[:, 0], tr[:, 1])
    loss = model.criterion(pred, labels)
    loss.backward()
    hook.remove()
    all_ent, sub_emb, rel_emb = seen["inp"]
    logits = torch.mm(sub_emb * rel_emb, all_ent.t())      # operations_lp.py:121-126 before the sigmoid
    assert torch.equal(torch.sigmoid(logits), pred.detach())
    print("config_c1: real reference fwd+bwd %.1f s, loss %.8f, max|logit| %.2f" %
          (time.time() - t0, loss.item(), float(logits.abs().max())))
    state1 = _sd(model)
    args32 = list(g.arg_trace)
    # "truth": the same REAL modules in float64 from the same initial state, with the reference's fp32 scoring head
    # (fp32 sigmoid + BCELoss on the logits rounded to fp32) -- in fp64 the probabilities would not saturate and
    # BCELoss's clamp / zero-gradient semantics, which ARE the reference's behaviour at this init, would vanish
    import copy
    model64 = copy.deepcopy(model)
    model64.load_state_dict(state0)
    model64 = model64.double().train()
    model64.zero_grad()

    class _Head32(nn.Module):
        def forward(self, all_ent, sub_emb, rel_emb):
            self.logits = torch.mm(sub_emb * rel_emb, all_ent.t())
            return torch.sigmoid(self.logits.float())
    model64.score_func = _Head32()
    g.edata['norm'] = g.edata['norm'].double()
    t0 = time.time()
    pred64 = model64(g, tr[:, 0], tr[:, 1])
    loss64 = model.criterion(pred64, labels)
    loss64.backward()
    g.edata['norm'] = g.edata['norm'].float()
    print("config_c1: fp64 truth %.1f s, loss %.8f" % (time.time() - t0, loss64.item()))
    out = {"dims": {"N": N, "R": R, "T": T, "D": D, "B": B}, "genotype": README_GENOTYPE,
           "truth64": {"loss": loss64.detach().clone(), "logits": summarize("logits", model64.score_func.logits),
                       "grads": _summ_all("grad.", _grads(model64)), "buffers": _buffers(model64)},
           "inputs": {"triples": checksum(trip), "subj": tr[:, 0].clone(), "rel": tr[:, 1].clone(),
                      "labels": summarize("labels", labels)},
           "loss": loss.detach().clone(), "pred": summarize("pred", pred), "logits": summarize("logits", logits),
           "all_ent": summarize("all_ent", all_ent),
           "saturated": {"p_eq_1": int((pred == 1).sum()), "p_eq_0": int((pred == 0).sum()), "numel": pred.numel()},
           "grads": _summ_all("grad.", _grads(model)),
           "buffers": {k: v.clone() for k, v in state1.items() if "running" in k or "num_batches" in k},
           "arg_vals": [a.reshape(-1)[positions(a.numel(), 1234 + i, 8192)].clone() for i, a in enumerate(args32)]}
    torch.save(out, os.path.join(OUT, "config_c1.pt"))


def gen_config_c3():
    """configs[2]/C3: LP supernet step (model_search_lp.py:131-194, search/mr_lp_search.py:188-236) on the
    WN18RR-shaped KG: N=40,943 R=11 T=86,835, D=200, init 100, num_base_r=23, 2 layers, zero/first/last nodes
    1/2/2, graph_batch_size 30,000, split 0.5, negative_sample 10, uniform edge sampler, np seed 0."""
    import time
    import utils.utils_rgcn as ref_rgcn
    from oracle.summary import checksum, summarize
    N, R, T, D, D0, GB = 40943, 11, 86835, 200, 100, 30000
    trip = synth_kg(N, R, T, seed=0)
    torch.manual_seed(0)
    np.random.seed(0)
    g, node_id, src_in, edge_type, node_norm, data, labels = ref_rgcn.generate_sampled_graph_and_labels(
        trip, GB, 0.5, R, None, None, 10, "uniform")
    node_id_t = torch.from_numpy(node_id).view(-1, 1).long()
    src_in_t, edge_type_t = torch.from_numpy(src_in), torch.from_numpy(edge_type)
    g.ndata['norm'] = torch.from_numpy(node_norm).view(-1, 1)                      # node_norm_to_edge_norm :30-36
    g.apply_edges(lambda edges: {'norm': edges.dst['norm'] * edges.src['norm']})
    data_t, labels_t = torch.from_numpy(data), torch.from_numpy(labels)
    model = ref_search_lp.Network('cpu', N, R, 2, 1, 2, 2, D, D0, 2 * R + 1, 40, 0.0, 0.0)
    model.apply(weights_init)
    model.train()
    state0 = _sd(model)
    alphas0 = [a.detach().clone() for a in model.arch_parameters()]
    t0 = time.time()
    ent_embed, rel_embed = model(g, node_id_t, src_in_t, edge_type_t)
    loss = model.get_loss(g, ent_embed, rel_embed, data_t, labels_t)
    loss.backward()
    print("config_c3: real reference supernet fwd+bwd %.1f s, loss %.8f, %d nodes, %d edges, %d samples" %
          (time.time() - t0, loss.item(), len(node_id), len(src_in), len(data)))
    _, dst, _ = g.edges(form='all')
    import copy
    model64 = copy.deepcopy(model)
    model64.load_state_dict(state0)
    model64 = model64.double().train()
    model64.zero_grad()
    for a, a0 in zip(model64.arch_parameters(), alphas0):
        a.data.copy_(a0)
        a.grad = None
    g.edata['norm'] = g.edata['norm'].double()
    # MixedOp casts every candidate output with .float() (cell_lp.py:30): make that a no-op for the fp64 run only
    _float = torch.Tensor.float
    torch.Tensor.float = lambda self, *a, **k: self
    try:
        e64, r64 = model64(g, node_id_t, src_in_t, edge_type_t)
        loss64 = model64.get_loss(g, e64, r64, data_t, labels_t.double())
        loss64.backward()
    finally:
        torch.Tensor.float = _float
    g.edata['norm'] = g.edata['norm'].float()
    print("config_c3: fp64 truth loss %.10f" % loss64.item())
    out = {"dims": {"N": N, "R": R, "T": T, "D": D, "D0": D0, "graph_batch_size": GB, "negative_sample": 10},
           "truth64": {"loss": loss64.detach().clone(), "ent_embed": summarize("ent_embed", e64),
                       "grads": _summ_all("grad.", _grads(model64)), "buffers": _buffers(model64),
                       "dalphas": [a.grad.clone() if a.grad is not None else None for a in model64.arch_parameters()]},
           "inputs": {"triples": checksum(trip), "node_id": checksum(node_id), "src": checksum(src_in),
                      "dst": checksum(dst.numpy()), "etype": checksum(edge_type), "samples": checksum(data),
                      "norm": summarize("norm", g.edata['norm'])},
           "loss": loss.detach().clone(), "ent_embed": summarize("ent_embed", ent_embed),
           "rel_embed": summarize("rel_embed", rel_embed), "grads": _summ_all("grad.", _grads(model)),
           "dalphas": [a.grad.clone() if a.grad is not None else None for a in model.arch_parameters()],
           "genotypes": str(model.show_genotypes()),
           "buffers": {k: v.clone() for k, v in _sd(model).items() if "running" in k or "num_batches" in k}}
    torch.save(out, os.path.join(OUT, "config_c3.pt"))


def gen_config_c2():
    """configs[1]/C2: NC derived network (default genotype, op_norm) on the AIFB-shaped graph: 8,285 nodes, 90 edge
    types, 58,086 directed edges, 4 classes, D=64, init 16, num_base_r=50, 2 layers, one 64-seed mini-batch of
    2-layer full-neighbour blocks (mr_nc_train.py:42-72, models/model.py:152-199)."""
    import time
    import models.model as ref_model_nc
    from collections import namedtuple
    from oracle.mrg_oracle import synth_nc_graph
    from oracle.summary import checksum, summarize
    G2 = namedtuple('Genotype', 'alpha_cell concat_node score_func', defaults=(None,))
    N, ET, E, D, D0, C, NB, B = 8285, 90, 58086, 64, 16, 4, 50, 64
    gr = synth_nc_graph(N, ET, E, C, 176, seed=0)
    src, dst, etype = gr["src"], gr["dst"], gr["etype"]
    seeds = np.sort(gr["labelled"][:B])
    blocks = _nc_blocks(src, dst, etype, seeds, 2)
    trip_index = torch.from_numpy(np.stack([np.arange(E), src, dst], 1)).long()
    labels = torch.from_numpy(gr["labels"][seeds]).long()
    torch.manual_seed(0)
    args = types.SimpleNamespace(feature_dim=D, op_norm=True)
    model = ref_model_nc.Network('cpu', eval(NC_GENOTYPE, {"Genotype": G2}), N, C, ET, 2, 1, 2, D, D0, NB,
                                 nn.CrossEntropyLoss(), args)
    model.apply(weights_init)
    model.train()
    state0 = _sd(model)
    t0 = time.time()
    logits = model(trip_index, [b for (b, _, _, _) in blocks])
    loss = nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    print("config_c2: real reference NC fwd+bwd %.1f s, loss %.8f, block edges %s" %
          (time.time() - t0, loss.item(), [len(e) for (_, e, _, _) in blocks]))
    import copy
    model64 = copy.deepcopy(model)
    model64.load_state_dict(state0)
    model64 = model64.double().train()
    model64.zero_grad()
    logits64 = model64(trip_index, [b for (b, _, _, _) in blocks])
    loss64 = nn.CrossEntropyLoss()(logits64, labels)
    loss64.backward()
    out = {"dims": {"N": N, "ET": ET, "E": E, "D": D, "D0": D0, "C": C, "NB": NB, "B": B}, "genotype": NC_GENOTYPE,
           "truth64": {"loss": loss64.detach().clone(), "logits": logits64.detach().clone(),
                       "grads": _summ_all("grad.", _grads(model64)), "buffers": _buffers(model64)},
           "inputs": {"src": checksum(src), "dst": checksum(dst), "etype": checksum(etype), "seeds": torch.from_numpy(seeds),
                      "block_eids": [checksum(e) for (_, e, _, _) in blocks]},
           "loss": loss.detach().clone(), "logits": logits.detach().clone(), "grads": _summ_all("grad.", _grads(model)),
           "buffers": {k: v.clone() for k, v in _sd(model).items() if "running" in k or "num_batches" in k}}
    torch.save(out, os.path.join(OUT, "config_c2.pt"))


def gen_score_transe():
    """sf_TransE_op (operations_lp.py:101-112) from the REAL reference: probabilities, BCE loss and gradients for
    two shapes (ragged tile edges; exact-zero differences so that sign(0) = 0 matters), fp32 + fp64."""
    cases = {}
    for tag, (B, N, D, gamma, seed) in {"small": (5, 37, 8, 9.0, 3), "ragged": (70, 333, 72, 40.0, 4)}.items():
        torch.manual_seed(seed)
        op = ref_lp.MIXED_OPS_sf['sf_TransE']({'gamma': gamma})
        ent = torch.randn(N, D)
        sub, rel = torch.randn(B, D), torch.randn(B, D)
        ent[3] = sub[1] + rel[1]                     # a zero distance: every |.| has a zero argument
        ent[5, :4] = (sub[2] + rel[2])[:4]
        label = (torch.rand(B, N) < 0.05).float() * 0.9 + 1.0 / N
        out = {}
        for dt in (torch.float32, torch.float64):
            e, s_, r_ = (t.detach().clone().to(dt).requires_grad_(True) for t in (ent, sub, rel))
            pred = op(e, s_, r_)
            loss = nn.BCELoss()(pred, label.to(dt))
            loss.backward()
            out[dt] = {"pred": pred.detach(), "loss": loss.detach(), "dent": e.grad.clone(), "dsub": s_.grad.clone(),
                       "drel": r_.grad.clone()}
        cases[tag] = {"B": B, "N": N, "D": D, "gamma": gamma, "ent": ent, "sub": sub, "rel": rel, "label": label,
                      "f32": out[torch.float32], "f64": out[torch.float64]}
    torch.save(cases, os.path.join(OUT, "score_transe.pt"))


def gen_labels():
    """1-N training items and their dense smoothed label rows from the REAL process() (utils/process_data.py:4-31)
    and TrainDataset (utils/data_set.py:6-33): the fixture behind the host-side batch builders and the device-side
    label expansion."""
    N, R, T = 311, 4, 900
    trip = synth_kg(N, R, T, seed=5)
    items = process({'train': trip, 'valid': trip[:0], 'test': trip[:0]}, R)['train'][:24]
    out = {"N": N, "R": R, "triples": torch.from_numpy(trip), "items": items, "rows": {}}
    for ls in (0.1, 0.0):
        ds = TrainDataset(items, N, types.SimpleNamespace(lbl_smooth=ls))
        trs, ys = zip(*[ds[i] for i in range(len(items))])
        out["rows"][ls] = (torch.stack(trs), torch.stack(ys))
    torch.save(out, os.path.join(OUT, "labels.pt"))
    print("labels:", len(items), "items")


def gen_predict():
    """The REAL predict() of train/mr_lp_train.py:269-314 (imported from the script with stub modules for its
    non-path imports) on pre-computed probability matrices: a fake model returns them batch by batch.  Scores are
    distinct inside every row, so the (upstream unspecified) order of exact ties plays no role."""
    import importlib.util
    for name, attrs in (("tensorboardX", {"SummaryWriter": object}), ("dataloader", {"get_dataset": None}),
                        ("dgl.contrib", {}), ("dgl.contrib.data", {"load_data": None})):
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules.setdefault(name, mod)
    spec = importlib.util.spec_from_file_location("ref_lp_train", os.path.join(REF, "train", "mr_lp_train.py"))
    ref_train = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_train)
    torch.manual_seed(11)
    rng = np.random.RandomState(11)
    N, B, nb = 523, 17, 5
    batches = []
    for _ in range(nb):
        pred = torch.rand(B, N) * 0.98 + 0.01              # distinct with probability 1
        trip = torch.from_numpy(np.stack([rng.randint(0, N, B), rng.randint(0, 6, B), rng.randint(0, N, B)], 1))
        lab = (torch.rand(B, N) < 0.03).float()
        lab[torch.arange(B), trip[:, 2]] = 1.0
        batches.append((pred, trip, lab))

    class FakeModel:
        def __init__(self):
            self.k = 0

        def eval(self):
            return self

        def __call__(self, g, subj, rel):
            out = batches[self.k][0].clone()
            self.k += 1
            return out

    results, loss = ref_train.predict([(t, l) for _, t, l in batches], None, FakeModel(), "cpu")
    torch.save({"batches": batches, "results": results, "loss": float(loss)}, os.path.join(OUT, "predict.pt"))
    print("predict:", results)


if __name__ == "__main__":
    single = {"predict": gen_predict, "labels": gen_labels, "score_transe": gen_score_transe, "config_c1": gen_config_c1, "config_c2": gen_config_c2,
              "config_c3": gen_config_c3}
    if len(sys.argv) > 1 and sys.argv[1] in single:
        for name in sys.argv[1:]:
            single[name]()
        sys.exit(0)
    os.makedirs(OUT, exist_ok=True)
    gen_ops_lp()
    gen_ops_nc()
    gen_network_lp()
    gen_mixed_op()
    gen_search_lp()
    gen_compgcn()
    gen_network_nc()
    gen_predict()
    gen_labels()
    gen_score_transe()
    gen_config_c2()
    gen_config_c3()
    gen_config_c1()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
