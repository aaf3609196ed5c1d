"""CPU: the oracle restatement (oracle/mrg_oracle.py) against golden vectors produced by
the REAL reference code (oracle/make_golden.py).  Pins the oracle."""
import os
from collections import namedtuple

import numpy as np
import pytest
import torch

from oracle import mrg_oracle as O

Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func")
TOL = 2e-6


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def _close(a, b, tol=TOL):
    scale = max(1.0, float(b.abs().max()))
    err = float((a.detach() - b).abs().max())
    assert err <= tol * scale, f"max err {err:.3e} scale {scale:.3e}"


def test_build_graph_bit_exact(golden_dir):
    for f in ["ops_lp.pt", "network_lp.pt", "mixed_op.pt"]:
        gd = _load(golden_dir, f)["graph"]
        g = O.build_graph(gd["num_ent"], gd["triples"].numpy(), gd["num_rels"])
        assert np.array_equal(g["src"], gd["src"].numpy())
        assert np.array_equal(g["dst"], gd["dst"].numpy())
        assert np.array_equal(g["etype"], gd["etype"].numpy())
        assert np.array_equal(g["in_deg"], gd["in_deg"].numpy())
        assert np.array_equal(g["norm"], gd["norm"].numpy())  # same numpy float32 ops -> bit exact


LP_OPS = ['pre_mult', 'pre_sub', 'pre_add', 'f_zero', 'f_identity', 'f_dense', 'f_dense_comp', 'f_comp',
          'f_sparse', 'f_sparse_comp', 'f_dense_last', 'f_sparse_last', 'a_max', 'a_mean', 'a_sum']


@pytest.mark.parametrize("name", LP_OPS)
def test_lp_ops(golden_dir, name):
    G = _load(golden_dir, "ops_lp.pt")
    gd, c = G["graph"], G["cases"][name]
    E, N = gd["src"].numel(), gd["num_ent"]
    x = c["x"].clone().requires_grad_(True)
    xin = c["xin"].clone().requires_grad_(True)
    P = {"op." + k: v.clone().requires_grad_(True) for k, v in c["state"].items()}
    out = O.apply_op_lp(name, P, "op", x, xin, E, gd["norm"], gd["dst"], N)
    _close(out, c["out"])
    if name == "a_max":
        _, arg = O.a_op_lp(name, P, "op", x, E, gd["dst"], N, return_arg=True)
        assert torch.equal(arg, c["arg"])
    if out.requires_grad:
        out.backward(c["cot"])
        if c["dx"] is not None:
            _close(x.grad, c["dx"])
        if c["dxin"] is not None and xin.grad is not None:
            _close(xin.grad, c["dxin"])
        for k, g in c["dparams"].items():
            if g is not None:
                _close(P["op." + k].grad, g)


@pytest.mark.parametrize("name", ['a_max', 'a_mean', 'a_sum', 'a_std'])
def test_nc_aggregators(golden_dir, name):
    G = _load(golden_dir, "ops_nc.pt")
    c = G["cases"][name]
    x = c["x"].clone().requires_grad_(True)
    P = {"op." + k: v.clone().requires_grad_(True) for k, v in c["state"].items()}
    out = O.a_op_nc(name, P, "op", x, G["dst"], G["n_dst"])
    _close(out, c["out"])
    out.backward(c["cot"])
    _close(x.grad, c["dx"], tol=1e-5)
    for k, g in c["dparams"].items():
        _close(P["op." + k].grad, g, tol=1e-5)


def test_network_lp_forward_backward(golden_dir):
    G = _load(golden_dir, "network_lp.pt")
    gd = G["graph"]
    genos = eval(G["genotype"])
    P = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in G["state0"].items()}
    graph = O.build_graph(gd["num_ent"], gd["triples"].numpy(), gd["num_rels"])
    pred = O.network_lp(genos, P, graph, G["subj"], G["rel"], gd["num_rels"], training=True)
    _close(pred, G["pred"])
    loss = O.bce_loss(pred, G["labels"])
    _close(loss, G["loss"])
    loss.backward()
    for k, g in G["grads"].items():
        if g is not None:
            _close(P[k].grad, g, tol=1e-5)
    # eval mode with the reference's running stats after its 4 training steps
    Pe = G["state_eval"]
    pe = O.network_lp(genos, Pe, graph, G["subj"], G["rel"], gd["num_rels"], training=False)
    _close(pe, G["pred_eval"])


def test_labels_match_dataset(golden_dir):
    G = _load(golden_dir, "network_lp.pt")
    gd = G["graph"]
    items = O.process_1n(gd["triples"].numpy(), gd["num_rels"])[: G["dims"]["B"]]
    for it, (tr, lab) in zip(items, G["train_items"]):
        assert list(it["triple"]) == tr and sorted(lab) == it["label"]
    y = O.smoothed_labels(items, gd["num_ent"], 0.1)
    assert torch.equal(y, G["labels"])


@pytest.mark.parametrize("tag", ["pre", "first", "middle", "last"])
def test_mixed_op(golden_dir, tag):
    G = _load(golden_dir, "mixed_op.pt")
    gd, c = G["graph"], G["cases"][tag]
    E, N = gd["src"].numel(), gd["num_ent"]
    alpha = c["alpha"].clone().requires_grad_(True)
    x = c["x"].clone().requires_grad_(True)
    xin = c["xin"].clone().requires_grad_(True)
    P = {"mo." + k: v.clone().requires_grad_(v.is_floating_point()) for k, v in c["state"].items()}
    out = O.mixed_op_lp(c["names"], torch.softmax(alpha, 0), P, "mo", x, xin, E, gd["norm"], gd["dst"], N)
    _close(out, c["out"])
    out.backward(c["cot"])
    _close(alpha.grad, c["dalpha"], tol=1e-5)
    _close(x.grad, c["dx"], tol=1e-5)


@pytest.mark.parametrize("comp", ["sub", "mul", "ccorr"])
def test_compgcn(golden_dir, comp):
    G = _load(golden_dir, "compgcn.pt")
    gd, c = G["graph"], G["cases"][comp]
    E = gd["src"].numel()
    h = c["h"].clone().requires_grad_(True)
    r = c["r"].clone().requires_grad_(True)
    P = {"l." + k: v.clone().requires_grad_(v.is_floating_point()) for k, v in c["state"].items()}
    in_mask = torch.arange(E) < E // 2
    n_out, r_out = O.comp_graph_conv(P, "l", comp, h, r, gd["src"], gd["dst"], gd["etype"], gd["norm"].view(-1),
                                     in_mask, ~in_mask)
    _close(n_out, c["n_out"])
    _close(r_out, c["r_out"])
    ((n_out * c["c1"]).sum() + (r_out * c["c2"]).sum()).backward()
    _close(h.grad, c["dh"], tol=1e-5)
    _close(r.grad, c["dr"], tol=1e-5)
    for k, g in c["dparams"].items():
        if g is not None:
            _close(P["l." + k].grad, g, tol=1e-5)


def test_oracle_predict_matches_real_reference_predict(golden_dir):
    """Evaluation path: the oracle's sort-free filtered rank reproduces the result dictionary and
    the summed test loss of the REAL predict() (train/mr_lp_train.py:269-314), fixture tests/golden/predict.pt."""
    import os
    import torch
    from oracle import mrg_oracle as O
    G = torch.load(os.path.join(golden_dir, "predict.pt"), weights_only=False)
    results, loss = O.predict_results(G["batches"])
    assert results == G["results"], (results, G["results"])
    assert abs(loss - G["loss"]) <= 1e-6 * max(1.0, abs(G["loss"]))
