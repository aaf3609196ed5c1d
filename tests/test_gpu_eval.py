"""Evaluation path: mr_gnas_b200.evaluate.predict against a literal restatement of the
reference's predict() (train/mr_lp_train.py:269-314: model(g, subj, rel) per batch, torch.where filter, double
argsort) on the same model, graph and batches.  Ranks are integers: the result dictionaries must be identical."""
import types
from collections import namedtuple

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func")
README = [Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_sparse_comp', 2, 1), ('f_sparse_comp', 3, 2), ('a_max', 4, 2),
                               ('a_max', 5, 3), ('f_sparse_last', 6, 5), ('f_sparse_last', 7, 5)],
                   concat_node=[4, 5, 6, 7], score_func='sf_DisMult')]


def _reference_predict(loader, g, model, device):
    """train/mr_lp_train.py:280-312, verbatim semantics (stable sorts so ties are defined)."""
    with torch.no_grad():
        results, test_loss = dict(), []
        model.eval()
        for triplets, labels in loader:
            triplets, labels = triplets.to(device), labels.to(device)
            subj, rel, obj = triplets[:, 0], triplets[:, 1], triplets[:, 2]
            pred = model(g, subj, rel)
            test_loss.append(F.binary_cross_entropy(pred, labels).item())
            b_range = torch.arange(pred.shape[0], device=device)
            target_pred = pred[b_range, obj]
            pred = torch.where(labels.byte().bool(), -torch.ones_like(pred) * 10000000, pred)
            pred[b_range, obj] = target_pred
            order = torch.sort(pred, dim=1, descending=True, stable=True).indices
            ranks = 1 + torch.sort(order, dim=1, descending=False, stable=True).indices[b_range, obj]
            ranks = ranks.float()
            results['count'] = torch.numel(ranks) + results.get('count', 0)
            results['mr'] = torch.sum(ranks).item() + results.get('mr', 0)
            results['mrr'] = torch.sum(1.0 / ranks).item() + results.get('mrr', 0)
            for k in [1, 3, 10]:
                results[f'hits@{k}'] = torch.numel(ranks[ranks <= k]) + results.get(f'hits@{k}', 0)
        return results, np.sum(test_loss)


def test_filtered_rank_kernel_with_ties():
    from mr_gnas_b200.evaluate import filtered_rank
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B, N = 37, 1001
    pred = torch.rand(B, N, device=dev)
    pred[:, ::7] = 0.5                       # many exact ties
    pred[3] = 1.0                            # a fully saturated row
    pred[4] = 0.0
    labels = (torch.rand(B, N, device=dev) < 0.1).float()
    obj = torch.randint(0, N, (B,), device=dev)
    obj[5] = 0
    obj[6] = N - 1
    labels[torch.arange(B), obj] = 1.0       # the target is always a known object
    got = filtered_rank(pred, labels, obj)
    p = torch.where(labels.bool(), torch.full_like(pred, -10000000.0), pred)
    p[torch.arange(B), obj] = pred[torch.arange(B), obj]
    order = torch.sort(p, dim=1, descending=True, stable=True).indices
    ref = 1 + torch.sort(order, dim=1, stable=True).indices[torch.arange(B), obj]
    assert torch.equal(got, ref)


def test_predict_matches_reference_loop():
    from mr_gnas_b200 import evaluate
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.model_lp import Network
    from mr_gnas_b200.synth import synth_kg
    from mr_gnas_b200.utils import weights_init
    dev = torch.device("cuda:0")
    N, R, T, D, B = 900, 7, 7000, 64, 50
    trip = synth_kg(N, R, T, seed=4)
    g = MRGraph.from_triples(N, trip, R, device=dev)
    args = types.SimpleNamespace(feature_dim=D, drop_aggr=0.0, drop_op=0.0, gamma=40, embed_dim=D, conve_hid_drop=0.0,
                                 feat_drop=0.0, num_filt=4, ker_sz=3, k_w=4, k_h=D // 4)
    torch.manual_seed(0)
    model = Network(dev, README, N, R, D, D, 2 * R + 1, nn.BCELoss(), 0.0, args)
    model.apply(weights_init)
    model = model.to(dev)
    # a few training steps so the running statistics are not the initial ones
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    rng = np.random.RandomState(1)
    for _ in range(3):
        subj = torch.from_numpy(rng.randint(0, N, 32)).to(dev)
        rel = torch.from_numpy(rng.randint(0, 2 * R, 32)).to(dev)
        lab = (torch.from_numpy(rng.rand(32, N)) < 0.02).float().to(dev)
        opt.zero_grad()
        model._loss(g, subj, rel, lab * 0.9 + 1.0 / N).backward()
        opt.step()
    # evaluation batches: (s, r, o) triples with multi-hot labels that contain o
    loader = []
    for _ in range(4):
        t = torch.from_numpy(np.stack([rng.randint(0, N, B), rng.randint(0, 2 * R, B), rng.randint(0, N, B)], 1))
        lab = (torch.from_numpy(rng.rand(B, N)) < 0.01).float()
        lab[torch.arange(B), t[:, 2]] = 1.0
        loader.append((t, lab))
    ref, ref_loss = _reference_predict(loader, g, model, dev)
    got, got_loss = evaluate.predict(loader, g, model, dev)
    assert got == ref, (got, ref)
    assert abs(got_loss - ref_loss) <= 1e-5 * max(1.0, abs(ref_loss))
    res, loss = evaluate.infer(model, g, loader, loader, dev)
    assert res['mr'] == round(ref['mr'] / ref['count'], 5) and res['left_mrr'] == res['right_mrr']


@pytest.mark.parametrize("ls", [0.1, 0.0])
def test_labels_from_csr_bit_identical_to_dense_host_labels(ls):
    """labels expanded on the device from the sparse object lists == TrainDataset's dense rows."""
    from mr_gnas_b200.process_data import labels_on_device, make_batch, make_batch_sparse, process
    from mr_gnas_b200.synth import synth_kg
    dev = torch.device("cuda:0")
    N, R, T, B = 14541, 11, 20000, 96
    trip = synth_kg(N, R, T, seed=2)
    items = process({'train': trip.tolist(), 'valid': [], 'test': []}, R)['train'][:B]
    t_dense, y = make_batch(items, N, lbl_smooth=ls)
    t_sp, ptr, idx = make_batch_sparse(items)
    assert torch.equal(t_dense, t_sp)
    out = labels_on_device(ptr.to(dev), idx.to(dev), N, ls)
    assert torch.equal(out.cpu(), y)
    lo, hi = 5000, 9000     # one destination-partition rank's columns
    part = labels_on_device(ptr.to(dev), idx.to(dev), N, ls, lo, hi)
    assert torch.equal(part.cpu(), y[:, lo:hi])
    assert idx.numel() * 4 + ptr.numel() * 4 < y.numel() * 4 / 100    # > 100x fewer bytes over PCIe


def test_filtered_rank_kernel_matches_real_reference_golden(golden_dir):
    """mrg_filtered_rank on the fixture generated by the REAL reference predict() (tests/golden/predict.pt)."""
    import os
    from mr_gnas_b200.evaluate import filtered_rank
    dev = torch.device("cuda:0")
    G = torch.load(os.path.join(golden_dir, "predict.pt"), weights_only=False)
    mr = mrr = 0.0
    hits = {1: 0, 3: 0, 10: 0}
    for pred, trip, lab in G["batches"]:
        ranks = filtered_rank(pred.to(dev), lab.to(dev), trip[:, 2].to(dev)).float()
        mr += torch.sum(ranks).item()
        mrr += torch.sum(1.0 / ranks).item()
        for k in hits:
            hits[k] += torch.numel(ranks[ranks <= k])
    ref = G["results"]
    assert mr == ref["mr"] and abs(mrr - ref["mrr"]) <= 1e-6 and all(hits[k] == ref[f"hits@{k}"] for k in hits)


@pytest.mark.parametrize("tag", ["small", "ragged"])
def test_transe_scorer_matches_real_reference_golden(golden_dir, tag):
    """sf_TransE: fused L1-distance kernels (mrg_transe_fwd / mrg_transe_bwd) against the REAL
    reference's sf_TransE_op (operations_lp.py:101-112) -- probabilities, BCE loss, gradients of the entity table,
    subject and relation rows; includes exact-zero differences (sign(0) = 0) and ragged tile edges.
    Bars vs the fp64 run of the same module: max(1e-5, 4 x the fp32 reference's own error)."""
    import os
    import torch.nn as nn
    from mr_gnas_b200.operations_lp import MIXED_OPS_sf
    from parity import check_grads, rel_err
    c = torch.load(os.path.join(golden_dir, "score_transe.pt"), weights_only=False)[tag]
    dev = torch.device("cuda:0")
    op = MIXED_OPS_sf['sf_TransE']({'gamma': c["gamma"]})
    ent, sub, rel = (c[k].to(dev).requires_grad_(True) for k in ("ent", "sub", "rel"))
    label = c["label"].to(dev)
    pred = op(ent, sub, rel)
    r32, r64 = c["f32"], c["f64"]
    assert rel_err(pred, r64["pred"]) <= max(1e-5, 4 * rel_err(r32["pred"], r64["pred"]))
    loss = nn.BCELoss()(pred, label)
    assert abs(float(loss) - float(r64["loss"])) <= max(1e-5, 4 * abs(float(r32["loss"]) - float(r64["loss"])) /
                                                        float(r64["loss"])) * float(r64["loss"])
    loss.backward()
    ours = {"dent": ent.grad, "dsub": sub.grad, "drel": rel.grad}
    check_grads("sf_TransE " + tag, ours, {k: r32[k] for k in ours}, {k: r64[k] for k in ours})
    # the fused-loss entry point used by Network._loss gives the same loss and gradients
    ent2, sub2, rel2 = (c[k].to(dev).requires_grad_(True) for k in ("ent", "sub", "rel"))
    l2 = op.loss(ent2, sub2, rel2, label)
    assert abs(float(l2) - float(loss)) <= 1e-6 * float(loss)
    l2.backward()
    assert rel_err(ent2.grad, ent.grad) <= 1e-5 and rel_err(sub2.grad, sub.grad) <= 1e-5
