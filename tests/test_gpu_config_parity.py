"""GPU parity AT BASELINE.json's configuration shapes (nothing rescaled), against
  (a) the REAL reference run at that shape by oracle/make_golden.py (fp32 results + an fp64 "truth" run of the
      same reference modules; compact summaries in tests/golden/config_c{1,2,3}.pt), and
  (b) for C1, the CPU oracle restatement executed live on the same seeded inputs (full tensors).

Bars (tests/parity.py, tests/config_cases.py): every float tensor relative to its own scale; against the fp64
truth the product may be no further than max(1e-5, 4 x the real fp32 reference's own error); integer routing
(a_max argmax) >= 99.5 % identical to the reference (near-ties may resolve differently under any fp32 GEMM
re-ordering; exactness on tie-free data is pinned by the small goldens)."""
import types
from collections import namedtuple

import numpy as np
import pytest
import torch
import torch.nn as nn

from config_cases import (c1_inputs, c2_inputs, c3_inputs, check_vs_truth, load, loss_bar, lp_args)
from oracle import mrg_oracle as O
from oracle.summary import errors, positions, sample
from parity import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func", defaults=(None,))


from config_cases import assert_report as _report  # noqa: E402


def test_c1_full_train_step_vs_real_reference_and_oracle(golden_dir):
    """README genotype LP training step at C1: N=14,541 R=237 T=272,115 (E=544,230) D=200 B=256, label smoothing
    0.1, seeded init as mr_lp_train.py:94-113 draws it, w_rel NOT rescaled (max|logit| = 606, 4 % of the
    probabilities saturate).  reference: models/model_lp.py:123-150, train/mr_lp_train.py:222-246."""
    from mr_gnas_b200.functional import decode_arg
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.model_lp import Network
    from mr_gnas_b200.utils import weights_init
    G = load(golden_dir, "config_c1.pt")
    d, T64 = G["dims"], G["truth64"]
    trip, subj, rel, labels = c1_inputs(G)
    genos = eval(G["genotype"])
    torch.manual_seed(0)
    model = Network('cpu', genos, d["N"], d["R"], d["D"], d["D"], 2 * d["R"] + 1, nn.BCELoss(), 0.0, lp_args(d["D"]))
    model.apply(weights_init)
    P = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    model = model.to(DEV).train()
    g = MRGraph.from_triples(d["N"], trip, d["R"], device=DEV)
    g.arg_trace = []
    loss = model._loss(g, subj.to(DEV), rel.to(DEV), labels.to(DEV))
    loss.backward()
    args_gpu = [decode_arg(a).cpu() for a in g.arg_trace]
    g.arg_trace = None
    # BatchNorm running statistics after one training forward (reference buffers, fp32)
    for k, v in model.state_dict().items():
        if "running" in k:
            t64 = T64["buffers"][k]
            bar = max(1e-5, 4 * rel_err(G["buffers"][k], t64))
            assert rel_err(v, t64) <= bar, (k, rel_err(v, t64), bar)
        if "num_batches" in k:
            assert int(v) == int(G["buffers"][k]), k
    grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
    with torch.no_grad():           # the logits of the same step (second forward: buffers already checked)
        ent, rel_embed = model._embed(g)
        z = ((ent[subj.to(DEV)] * rel_embed[rel.to(DEV)]) @ ent.t()).cpu()
    # ---- (a) against the real reference (+ its fp64 truth)
    rep = []
    check_vs_truth("logits", z, G["logits"], T64["logits"], report=rep)
    tol, width, nb = loss_bar(z, sample("logits", z), G["logits"]["vals"], G["loss"])
    print(f"C1 loss: ours {float(loss):.8f}, real reference {float(G['loss']):.8f}, fp64-network truth "
          f"{float(T64['loss']):.8f}; {nb} logits within {width:.1e} of the saturation point -> bar {tol:.2e}")
    assert abs(float(loss) - float(G["loss"])) <= tol
    for k, summ in G["grads"].items():
        if summ is None:
            assert k not in grads
            continue
        check_vs_truth("grad." + k, grads[k], summ, T64["grads"][k], report=rep)
    _report("C1 vs real reference", rep)
    for i, a in enumerate(args_gpu[:2]):
        got = a.reshape(-1)[positions(a.numel(), 1234 + i, 8192)].long()
        agree = float((got == G["arg_vals"][i]).float().mean())
        print(f"a_max #{i}: argmax identical to the real reference on {agree:.5f} of 8192 sampled (node, feature) pairs")
        assert agree >= 0.995
    # ---- (b) against the oracle restatement executed now, full tensors
    O.ARG_TRACE = []
    try:
        pred_o, ent_o, rel_o = O.network_lp(genos, P, O.build_graph(d["N"], trip, d["R"]), subj, rel, d["R"],
                                            training=True, return_emb=True)
        args_o = O.ARG_TRACE
    finally:
        O.ARG_TRACE = None
    loss_o = O.bce_loss(pred_o, labels)
    loss_o.backward()
    with torch.no_grad():
        z_o = (ent_o[subj] * rel_o[rel]) @ ent_o.t()
    tol_o, width_o, nb_o = loss_bar(z_o, z_o, z, loss_o)
    assert abs(float(loss) - float(loss_o)) <= tol_o, (float(loss), float(loss_o), nb_o, tol_o)
    worst = []
    for k, summ in G["grads"].items():
        if summ is None:
            continue
        # |ours - oracle| <= |ours - truth| + |oracle - truth|: twice the per-tensor bar used above
        from config_cases import ref_error
        bar = 2.0 * max(1e-5, 4.0 * ref_error(summ, T64["grads"][k])[0])
        e = rel_err(grads[k], P[k].grad)
        worst.append((e / bar, e, bar, k))
        assert e <= bar, (k, e, bar)
    print("C1 vs live oracle (full tensors): loss ours %.8f oracle %.8f; worst grad rel err %.2e (bar %.2e) at %s" %
          (float(loss), float(loss_o), *max(worst)[1:]))
    for i, (a, b) in enumerate(zip(args_gpu[:2], args_o)):
        agree = float((a.long() == b).float().mean())
        print(f"a_max #{i}: argmax identical to the oracle on {agree:.6f} of all {a.numel()} (node, feature) pairs")
        assert agree >= 0.995


def test_c3_supernet_step_vs_real_reference(golden_dir):
    """LP supernet step (all candidates mixed by softmax alphas) at C3: WN18RR-shaped KG N=40,943 R=11 T=86,835,
    graph_batch_size 30,000 (22,619 sampled nodes, 30,000 directed graph edges, 330,000 scored triplets), D=200,
    init 100, 2 layers.  reference: models/model_search_lp.py:131-194, search/mr_lp_search.py:188-236."""
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.model_search_lp import Network
    from mr_gnas_b200.utils import weights_init
    G = load(golden_dir, "config_c3.pt")
    d, T64 = G["dims"], G["truth64"]
    s = c3_inputs(G)
    torch.manual_seed(0)
    np.random.seed(0)
    model = Network('cpu', d["N"], d["R"], 2, 1, 2, 2, d["D"], d["D0"], 2 * d["R"] + 1, 40, 0.0, 0.0)
    model.apply(weights_init)
    assert str(model.show_genotypes()) == G["genotypes"]
    alphas0 = [a.detach().clone() for a in model.arch_parameters()]
    model = model.to(DEV)
    model._device = DEV
    alphas = [a.to(DEV).requires_grad_(True) for a in alphas0]
    (model.alphas_zero_cell, model.alphas_first_cell, model.alphas_middle_cell, model.alphas_last_cell,
     model.alphas_final_cell) = alphas
    model._arch_parameters = alphas
    model.train()
    n = s["num_nodes"]
    g = MRGraph.from_edges(s["src"], s["dst"], s["etype"], n, 2 * d["R"] + 1, device=DEV)
    assert np.array_equal(g.edge_norm.cpu().numpy(), s["norm"])
    g.edata['norm'] = g.edge_norm.view(-1, 1)            # [E,1] as node_norm_to_edge_norm leaves it
    dev_t = lambda a: torch.from_numpy(np.asarray(a)).to(DEV)
    node_id = dev_t(s["uniq_v"]).view(-1, 1).long()
    ent, rel_embed = model(g, node_id, dev_t(s["src"]), dev_t(s["etype"]))
    loss = model.get_loss(g, ent, rel_embed, dev_t(s["samples"]), dev_t(s["labels"]))
    loss.backward()
    rep = []
    check_vs_truth("ent_embed", ent.detach().cpu(), G["ent_embed"], T64["ent_embed"], report=rep)
    e32 = abs(float(G["loss"]) - float(T64["loss"])) / abs(float(T64["loss"]))
    e_us = abs(float(loss) - float(T64["loss"])) / abs(float(T64["loss"]))
    print(f"C3 loss: ours {float(loss):.8f}, real reference {float(G['loss']):.8f}, fp64 truth {float(T64['loss']):.10f}")
    assert e_us <= max(1e-5, 4 * e32)
    for k, p in model.named_parameters():
        summ = G["grads"].get(k)
        if summ is None:
            continue
        check_vs_truth("grad." + k, p.grad.detach().cpu(), summ, T64["grads"][k], report=rep)
    _report("C3 vs real reference", rep)
    for a, r32, r64 in zip(alphas, G["dalphas"], T64["dalphas"]):
        if r32 is None:
            continue
        bar = max(1e-5, 4 * rel_err(r32, r64.float()))
        assert rel_err(a.grad.cpu(), r64.float()) <= bar, (rel_err(a.grad.cpu(), r64.float()), bar)
    for k, v in model.state_dict().items():
        if "running" in k:
            t64 = T64["buffers"][k]
            bar = max(1e-5, 4 * rel_err(G["buffers"][k], t64))
            assert rel_err(v, t64) <= bar, (k, rel_err(v, t64), bar)


def test_c2_nc_block_step_vs_real_reference(golden_dir):
    """NC derived network (default genotype, op_norm) at C2: AIFB-shaped graph, 8,285 nodes, 90 edge types, 58,086
    directed edges, D=64, one 64-seed batch of 2-layer full-neighbour blocks.  reference: models/model.py:152-199."""
    from mr_gnas_b200.graph import full_neighbor_blocks
    from mr_gnas_b200.model import Network
    from mr_gnas_b200.utils import weights_init
    from oracle.summary import checksum
    G = load(golden_dir, "config_c2.pt")
    d, T64 = G["dims"], G["truth64"]
    gr, seeds = c2_inputs(G)
    blocks = full_neighbor_blocks(gr["src"], gr["dst"], gr["etype"], seeds, 2, device=DEV)
    assert [checksum(b.edata['_ID'].cpu().numpy()) for b in blocks] == [tuple(c) for c in G["inputs"]["block_eids"]]
    torch.manual_seed(0)
    args = types.SimpleNamespace(feature_dim=d["D"], op_norm=True)
    model = Network('cpu', eval(G["genotype"]), d["N"], d["C"], d["ET"], 2, 1, 2, d["D"], d["D0"], d["NB"],
                    nn.CrossEntropyLoss(), args)
    model.apply(weights_init)
    model = model.to(DEV).train()
    model._device = DEV
    E = d["E"]
    trip_index = torch.from_numpy(np.stack([np.arange(E), gr["src"], gr["dst"]], 1)).long().to(DEV)
    labels = torch.from_numpy(gr["labels"][seeds]).long().to(DEV)
    logits = model(trip_index, blocks)
    loss = nn.CrossEntropyLoss()(logits, labels)
    loss.backward()
    bar = max(1e-5, 4 * rel_err(G["logits"], T64["logits"].float()))
    assert rel_err(logits.cpu(), T64["logits"].float()) <= bar, (rel_err(logits.cpu(), T64["logits"].float()), bar)
    e32 = abs(float(G["loss"]) - float(T64["loss"])) / abs(float(T64["loss"]))
    assert abs(float(loss) - float(T64["loss"])) / abs(float(T64["loss"])) <= max(1e-5, 4 * e32)
    rep = []
    for k, p in model.named_parameters():
        summ = G["grads"].get(k)
        if summ is None:
            continue
        check_vs_truth("grad." + k, p.grad.detach().cpu(), summ, T64["grads"][k], report=rep)
    _report("C2 vs real reference", rep)
    for k, v in model.state_dict().items():
        if "running" in k:
            t64 = T64["buffers"][k]
            bar = max(1e-5, 4 * rel_err(G["buffers"][k], t64))
            assert rel_err(v, t64) <= bar, (k, rel_err(v, t64), bar)


def test_loss_and_mrr_curves_follow_the_oracle_on_eighth_scale_c1():
    """North-star target "matching loss / MRR curves": 60 Adam steps (lr 1e-3, a new 256-query batch per step) of the
    README genotype on the 1/8-scale C1 graph (N=14,541 R=237 D=200, first 34,014 train triples) from the same seeded
    init, three ways: the CUDA path, the fp32 CPU oracle, and the oracle in float64 (with the reference's fp32
    sigmoid + BCELoss head, as in make_golden's truth runs).  Then the filtered MRR / MR / Hits@10 of each final
    model in eval mode (running statistics) on 1,024 training queries and 1,024 held-out ones -- CUDA through
    evaluate.predict (mrg_filtered_rank), oracle through its restatement of predict().
    reference: train/mr_lp_train.py:222-246, 269-314.

    Adam turns every rounding difference into an O(lr) parameter difference within a few steps (sign(g) updates on
    small gradients) and BCELoss is discontinuous where sigmoid saturates, so ANY two floating-point evaluations of
    this training run separate: the fp32 and fp64 CPU curves themselves do.  The bar is therefore relative to that
    separation: at every step the running-max distance between the CUDA curve and the fp32 oracle curve may be at
    most 8 x the running-max distance between the oracle's own fp32 and fp64 curves."""
    import torch.nn.functional as F
    from mr_gnas_b200.evaluate import predict
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.model_lp import Network
    from mr_gnas_b200.process_data import make_batch, process
    from mr_gnas_b200.utils import weights_init
    N, R, T, D, B, STEPS = 14541, 237, 272115 // 8, 200, 256, 60
    all_trip = O.synth_kg(N, R, 272115, seed=0)
    trip, valid = all_trip[:T], all_trip[T:T + 512]
    genos = eval("[Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_sparse_comp', 2, 1), ('f_sparse_comp', 3, 2), "
                 "('a_max', 4, 2), ('a_max', 5, 3), ('f_sparse_last', 6, 5), ('f_sparse_last', 7, 5)], "
                 "concat_node=[4, 5, 6, 7], score_func='sf_DisMult')]")
    data = process({'train': trip, 'valid': valid, 'test': valid[:0]}, R)
    items = data['train']
    rng = np.random.RandomState(7)
    batches = [make_batch([items[j] for j in rng.choice(len(items), B, replace=False)], N, lbl_smooth=0.1)
               for _ in range(STEPS)]
    ev_sets = {"train": data['train_tail'][:512] + data['train_head'][:512],
               "valid": data['valid_tail'] + data['valid_head']}
    ev_batches = {k: [make_batch(v[i:i + 256], N) for i in range(0, len(v), 256)] for k, v in ev_sets.items()}
    torch.manual_seed(0)
    model = Network('cpu', genos, N, R, D, D, 2 * R + 1, nn.BCELoss(), 0.0, lp_args(D))
    model.apply(weights_init)
    state0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    # ---- CUDA path
    model = model.to(DEV).train()
    g = MRGraph.from_triples(N, trip, R, device=DEV)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss_g = []
    for t, y in batches:
        opt.zero_grad()
        l = model._loss(g, t[:, 0].to(DEV), t[:, 1].to(DEV), y.to(DEV))
        l.backward()
        opt.step()
        loss_g.append(float(l))
        del l
    res_g = {k: predict(v, g, model, DEV)[0] for k, v in ev_batches.items()}
    # ---- oracle, fp32 and fp64
    graph = O.build_graph(N, trip, R)

    def oracle_run(dtype):
        P = {k: (v.clone().to(dtype) if v.is_floating_point() else v.clone()) for k, v in state0.items()}
        for v in P.values():
            if v.is_floating_point():
                v.requires_grad_(True)
        opt_o = torch.optim.Adam([v for v in P.values() if v.requires_grad], lr=1e-3)

        def probs(t, training):
            _, ent, rel_embed = O.network_lp(genos, P, graph, t[:, 0], t[:, 1], R, training=training, return_emb=True)
            z = torch.mm(ent[t[:, 0]] * rel_embed[t[:, 1]], ent.t())
            return torch.sigmoid(z.float())                  # the reference's fp32 scoring head
        losses = []
        O.TRACK_RUNNING_STATS = True
        try:
            for t, y in batches:
                opt_o.zero_grad()
                l = F.binary_cross_entropy(probs(t, True), y)
                l.backward()
                opt_o.step()
                losses.append(float(l))
        finally:
            O.TRACK_RUNNING_STATS = False
        with torch.no_grad():
            res = {k: O.predict_results([(probs(t, False), t, y) for t, y in v])[0] for k, v in ev_batches.items()}
        return losses, res

    loss_o, res_o = oracle_run(torch.float32)
    loss_t, res_t = oracle_run(torch.float64)
    d_gpu = np.maximum.accumulate([abs(a - b) / abs(b) for a, b in zip(loss_g, loss_o)])
    d_ref = np.maximum.accumulate([abs(a - b) / abs(b) for a, b in zip(loss_o, loss_t)])
    for k in list(range(0, STEPS, 6)) + [STEPS - 1]:
        print(f"step {k:3d}: loss ours {loss_g[k]:.6f} oracle fp32 {loss_o[k]:.6f} fp64 {loss_t[k]:.6f} | running-max rel "
              f"distance ours-oracle {d_gpu[k]:.2e}, oracle fp32-fp64 {d_ref[k]:.2e}")
    assert loss_g[-1] < 0.05 * loss_g[0], "the loss must fall"
    assert abs(loss_g[0] - loss_o[0]) <= 1e-4 * loss_o[0]            # step 0: before any chaos (saturation jumps only)
    ratio = max(dg / max(dr, 1e-6) for dg, dr in zip(d_gpu, d_ref))
    print(f"largest ratio of the two running-max distances over the {STEPS} steps: {ratio:.2f}")
    assert ratio <= 8.0
    for name in ev_sets:
        cnt = res_o[name]['count']
        assert res_g[name]['count'] == cnt
        row = {m: (res_g[name][m] / cnt, res_o[name][m] / cnt, res_t[name][m] / cnt) for m in ('mrr', 'mr', 'hits@10')}
        print(f"after {STEPS} steps, {name} queries (ours / oracle fp32 / oracle fp64): " +
              "; ".join(f"{m} {a:.5f} / {b:.5f} / {c:.5f}" for m, (a, b, c) in row.items()))
        # after 60 steps the model ranks barely better than chance (MRR ~3e-4 = a handful of lucky queries): the
        # mean rank is the stable statistic (2 %), MRR / Hits are held to 8 x the oracle's own fp32-fp64 gap or 25 %.
        # MRR and Hits@10 are COUNTING statistics over `cnt` (1,024) queries in which a single query weighs up to
        # 1 / cnt = 9.8e-4 -- three times the whole MRR here: one query entering the top ranks on one side (observed
        # after the weight gradients moved to 3xTF32: one validation query at rank 7, +1.4e-4) is not a parity
        # difference, so a quarter of one query's weight is the floor of their bar.
        for m, (a, b, c) in row.items():
            rel_bar = 0.02 if m == 'mr' else 0.25
            one_query = 0.0 if m == 'mr' else 0.25 / cnt
            assert abs(a - b) <= max(8.0 * abs(b - c), rel_bar * abs(b), one_query) + 1e-6, (name, m, a, b, c)
