"""GPU parity at model level: mr_gnas_b200.model_lp.Network (README genotype) vs golden vectors
from the real reference network and vs the oracle on a larger seeded graph."""
import os
import types
from collections import namedtuple

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import mrg_oracle as O

pytestmark = pytest.mark.gpu
Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func")
README = ("[Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_sparse_comp', 2, 1), ('f_sparse_comp', 3, 2), "
          "('a_max', 4, 2), ('a_max', 5, 3), ('f_sparse_last', 6, 5), ('f_sparse_last', 7, 5)], "
          "concat_node=[4, 5, 6, 7], score_func='sf_DisMult')]")


from parity import grad_errors, rel_err as _err  # max|a-b| / max|b|: relative to the tensor's own scale, no absolute floor


def _args(D):
    return types.SimpleNamespace(feature_dim=D, drop_aggr=0.0, drop_op=0.0, gamma=40, embed_dim=D,
                                 conve_hid_drop=0.0, feat_drop=0.0, num_filt=4, ker_sz=3, k_w=4, k_h=D // 4)


def _build(dev, genos, N, R, D, D0):
    from mr_gnas_b200.model_lp import Network
    return Network(dev, genos, N, R, D, D0, 2 * R + 1, nn.BCELoss(), 0.0, _args(D)).to(dev)


@pytest.fixture(params=[True, False], ids=["fused", "modular"])
def fused(request):
    from mr_gnas_b200 import model_lp
    old = model_lp.USE_FUSED_CELL
    model_lp.USE_FUSED_CELL = request.param
    yield request.param
    model_lp.USE_FUSED_CELL = old


def test_network_lp_golden(golden_dir, fused):
    dev = torch.device("cuda:0")
    G = torch.load(os.path.join(golden_dir, "network_lp.pt"), weights_only=False)
    gd, d = G["graph"], G["dims"]
    from mr_gnas_b200.graph import MRGraph
    g = MRGraph.from_triples(gd["num_ent"], gd["triples"].numpy(), gd["num_rels"], device=dev)
    model = _build(dev, eval(G["genotype"]), d["N"], d["R"], d["D"], d["D0"])
    model.load_state_dict(G["state0"])
    model.train()
    subj, rel, labels = G["subj"].to(dev), G["rel"].to(dev), G["labels"].to(dev)
    pred = model(g, subj, rel)
    assert _err(pred, G["pred"]) <= 1e-5
    loss = model.criterion(pred, labels)
    assert _err(loss, G["loss"]) <= 1e-5
    loss.backward()
    worst = max(grad_errors({k: p.grad for k, p in model.named_parameters()}, G["grads"]))
    # gradients of a 7-op cell with 9 BatchNorms, fp32 end to end: scale-relative 1e-5 on every tensor
    assert worst[0] <= 1e-5, worst
    # running statistics after one training forward
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            assert _err(v.float(), G["state1"][k].float()) <= 1e-5, k
    # fused loss path (_loss) == criterion(forward)
    model.zero_grad()
    model.load_state_dict(G["state0"])
    l2 = model._loss(g, subj, rel, labels)
    assert _err(l2, G["loss"]) <= 1e-5
    l2.backward()
    worst = max(grad_errors({k: p.grad for k, p in model.named_parameters()}, G["grads"]))
    assert worst[0] <= 1e-5, worst
    # short Adam loss curve (4 steps) against the reference's
    model.load_state_dict(G["state0"])
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        l = model._loss(g, subj, rel, labels)
        l.backward()
        opt.step()
        losses.append(l.item())
    np.testing.assert_allclose(losses, G["losses"], rtol=2e-4)
    # eval mode with the reference's final state
    model.load_state_dict(G["state_eval"])
    model.eval()
    with torch.no_grad():
        pe = model(g, subj, rel)
    assert _err(pe, G["pred_eval"]) <= 1e-5


OTHER = ("[Genotype(alpha_cell=[('pre_mult', 1, 0), ('f_sparse_comp', 2, 1), ('a_sum', 3, 2), ('a_max', 4, 1), "
         "('f_sparse_comp', 5, 2), ('a_max', 6, 5), ('f_sparse_last', 7, 3), ('f_dense_last', 8, 4)], "
         "concat_node=[3, 4, 6, 7, 8], score_func='sf_DisMult')]")


@pytest.mark.parametrize("D,geno", [(64, README), (200, README), (64, OTHER), (128, OTHER)], ids=["readme64", "readme200", "other64", "other128"])
def test_network_lp_oracle_seeded(D, geno, fused):
    README = geno  # noqa: F841 (shadows the module constant on purpose)
    dev = torch.device("cuda:0")
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.utils import weights_init
    N, R, T, D0, B = 2500, 12, 15000, 48, 32
    trip = O.synth_kg(N, R, T, seed=21)
    graph = O.build_graph(N, trip, R)
    genos = eval(geno)
    torch.manual_seed(3)
    model = _build('cpu', genos, N, R, D, D0)
    model.apply(weights_init)
    # keep the 1-N logits out of fp32 sigmoid saturation (|x| > 17): there BCELoss's -100 log clamp makes the
    # loss depend on the floating-point width itself, so an fp64 "truth" would not be comparable
    model.w_rel.data.mul_(0.1)
    P = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    items = O.process_1n(trip, R)[:B]
    subj = torch.tensor([it["triple"][0] for it in items])
    rel = torch.tensor([it["triple"][1] for it in items])
    labels = O.smoothed_labels(items, N, 0.1)
    pred_o = O.network_lp(genos, P, graph, subj, rel, R, training=True)
    loss_o = O.bce_loss(pred_o, labels)
    loss_o.backward()
    # fp64 oracle: measures the fp32 CPU path's own rounding error on this problem
    P64 = {k: (v.detach().double().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
    pred_64 = O.network_lp(genos, P64, graph, subj, rel, R, training=True)
    loss_64 = O.bce_loss(pred_64, labels.double())
    loss_64.backward()
    model = model.to(dev)
    model.train()
    g = MRGraph.from_triples(N, trip, R, device=dev)
    loss_g = model._loss(g, subj.to(dev), rel.to(dev), labels.to(dev))
    loss_g.backward()
    assert _err(loss_g, loss_o) <= 1e-5
    # vs the fp64 truth the GPU must be no further than the reference's own fp32 CPU path is (x4 slack)
    assert _err(loss_g, loss_64.float()) <= max(1e-5, 4 * _err(loss_o, loss_64.float()))
    rep = []
    for k, p in model.named_parameters():
        go, g64 = P[k].grad, P64[k].grad
        if go is None:
            continue
        e_gpu64, e_cpu64, e_gpu_cpu = _err(p.grad, g64.float()), _err(go, g64.float()), _err(p.grad, go)
        rep.append((e_gpu64, e_cpu64, e_gpu_cpu, k))
    worst = max(rep)
    print("worst grad err vs fp64: gpu %.2e cpu %.2e (gpu-vs-cpu %.2e) at %s" % worst)
    # the GPU path must be as close to the fp64 truth as the reference's own fp32 CPU path is (x4 slack),
    # or within 1e-5 of it in the scale-relative norm
    for e_gpu64, e_cpu64, e_gpu_cpu, k in rep:
        assert e_gpu64 <= max(1e-5, 4 * e_cpu64), (k, e_gpu64, e_cpu64)


def test_graphed_train_step_replays_the_eager_step():
    """train.GraphedTrainStep: replaying the captured CUDA graph must give exactly the losses and parameters of the
    same steps issued eagerly (same kernels, same order, deterministic reductions) -- dense and sparse-label inputs."""
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.process_data import make_batch, make_batch_sparse, process
    from mr_gnas_b200.train import GraphedTrainStep
    from mr_gnas_b200.utils import weights_init
    dev = torch.device("cuda:0")
    N, R, T, D, B = 700, 6, 6000, 64, 32
    trip = O.synth_kg(N, R, T, seed=9)
    g = MRGraph.from_triples(N, trip, R, device=dev)
    items = process({'train': trip.tolist(), 'valid': [], 'test': []}, R)['train']
    batches = [[items[j] for j in range(k * B, (k + 1) * B)] for k in range(4)]
    dense = [make_batch(b, N, lbl_smooth=0.1) for b in batches]
    sparse = [make_batch_sparse(b) for b in batches]

    def fresh():
        torch.manual_seed(0)
        m = _build('cpu', eval(README), N, R, D, D)
        m.apply(weights_init)
        m = m.to(dev).train()
        return m, torch.optim.Adam(m.parameters(), lr=1e-3, fused=True, capturable=True)

    # eager reference A: batches 0..3 from the fresh state (default capture() restores model + optimiser state
    # after its warm-up); eager reference B: 3 warm-up steps on batch 0 then batches 1..3 (keep_warmup_updates)
    t0, y0 = dense[0]
    m_a, opt_a = fresh()
    run_a = GraphedTrainStep(m_a, g, opt_a, B, N, warmup=3)
    losses_a = [float(run_a(t[:, 0].to(dev), t[:, 1].to(dev), y.to(dev))) for t, y in dense]
    m_e, opt_e = fresh()
    run_e = GraphedTrainStep(m_e, g, opt_e, B, N, warmup=3)
    for _ in range(3):
        run_e(t0[:, 0].to(dev), t0[:, 1].to(dev), y0.to(dev))
    losses_e = [float(run_e(t[:, 0].to(dev), t[:, 1].to(dev), y.to(dev))) for t, y in dense[1:]]
    for mode, keep in (("dense", False), ("dense", True), ("sparse", False), ("sparse", True)):
        m_g, opt_g = fresh()
        cfg = dict(num_ent=N, lbl_smooth=0.1, cap=4 * max(s[2].numel() for s in sparse)) if mode == "sparse" else None
        run_g = GraphedTrainStep(m_g, g, opt_g, B, N, warmup=3, sparse_labels=cfg)
        if mode == "dense":
            run_g.load(t0[:, 0].to(dev), t0[:, 1].to(dev), y0.to(dev))
        else:
            run_g.load_sparse(sparse[0][0][:, 0].to(dev), sparse[0][0][:, 1].to(dev), sparse[0][1].to(dev),
                              sparse[0][2].to(dev))
        run_g.capture(keep_warmup_updates=keep)
        assert run_g.graph is not None
        first = 1 if keep else 0
        if mode == "dense":
            losses_g = [float(run_g(t[:, 0].to(dev), t[:, 1].to(dev), y.to(dev))) for t, y in dense[first:]]
        else:
            losses_g = [float(run_g(t[:, 0].to(dev), t[:, 1].to(dev), label_csr=(p.to(dev), i.to(dev))))
                        for t, p, i in sparse[first:]]
        want, m_want = (losses_e, m_e) if keep else (losses_a, m_a)
        assert losses_g == want, (mode, keep, losses_g, want)
        for (k, a), b in zip(m_g.state_dict().items(), m_want.state_dict().values()):
            assert torch.equal(a, b), (mode, keep, k)
