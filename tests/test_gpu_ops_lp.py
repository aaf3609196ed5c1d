"""GPU parity: libmrgnas kernels (through the C ABI) vs the CPU oracle / golden vectors.

Tolerances: fp32 outputs and gradients within REL = 1e-5 of the oracle, measured relative to the reference
tensor's own scale, max|a-b| / max|b| (tests/parity.py; the north_star's "1e-5 relative in fp32" -- no
absolute floor, so small-magnitude gradients are held to the same relative bar); integer graph arrays and
argmax indices bit-exact (tie-break: lowest original edge id)."""
import os
from collections import namedtuple

import numpy as np
import pytest
import torch

from oracle import mrg_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-5
Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func")


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


from parity import rel_err as _err  # max|a-b| / max|b|: relative to the tensor's own scale, no absolute floor


def _check(name, a, b, tol=REL):
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    e = _err(a, b)
    assert e <= tol, f"{name}: rel err {e:.3e} > {tol:.1e}"


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _graph_from_golden(gd, dev):
    from mr_gnas_b200.graph import MRGraph
    return MRGraph.from_triples(gd["num_ent"], gd["triples"].numpy(), gd["num_rels"], device=dev)


# ------------------------------------------------------------------------------ K0
@pytest.mark.parametrize("N,R,T,seed", [(37, 4, 90, 3), (500, 7, 4000, 1), (2000, 11, 30000, 2), (64, 1, 1, 5)])
def test_graph_build_bit_exact(dev, N, R, T, seed):
    from mr_gnas_b200.graph import MRGraph
    trip = O.synth_kg(N, R, T, seed=seed)
    ref = O.build_graph(N, trip, R)
    g = MRGraph.from_triples(N, trip, R, device=dev)
    E = 2 * T
    assert np.array_equal(g.src.cpu().numpy(), ref["src"])
    assert np.array_equal(g.dst.cpu().numpy(), ref["dst"])
    assert np.array_equal(g.etype.cpu().numpy(), ref["etype"])
    assert np.array_equal(g.in_deg.cpu().numpy(), ref["in_deg"])
    ptr, eid = O.csr_by_dst(ref["dst"], N)
    assert np.array_equal(g.csr.ptr.cpu().numpy(), ptr)
    assert np.array_equal(g.csr.idx.cpu().numpy()[:E], eid)
    src_final = np.concatenate([ref["src"], np.arange(N)])
    ptr2, row2 = O.csr_by_dst(src_final, N)
    assert np.array_equal(g.csc.ptr.cpu().numpy(), ptr2)
    assert np.array_equal(g.csc.idx.cpu().numpy(), row2)
    et_final = np.concatenate([ref["etype"], np.full(N, 2 * R)])
    ptr3, row3 = O.csr_by_dst(et_final, 2 * R + 1)
    assert np.array_equal(g.rel.ptr.cpu().numpy(), ptr3)
    assert np.array_equal(g.rel.idx.cpu().numpy(), row3)
    # chunk tables cover every segment exactly
    cf = g.csr.chunk_first.cpu().numpy()
    assert np.array_equal(np.diff(cf), (np.diff(ptr) + 31) // 32)
    # degree norm: float32 in_deg**-0.5 products (mr_lp_train.py:82-86): bit-exact
    assert np.array_equal(g.edge_norm.cpu().numpy(), ref["norm"])


def test_graph_golden(dev, golden_dir):
    for f in ["ops_lp.pt", "network_lp.pt"]:
        gd = _load(golden_dir, f)["graph"]
        g = _graph_from_golden(gd, dev)
        assert torch.equal(g.src.cpu().long(), gd["src"])
        assert torch.equal(g.dst.cpu().long(), gd["dst"])
        assert torch.equal(g.etype.cpu().long(), gd["etype"])
        assert torch.equal(g.in_deg.cpu().long(), gd["in_deg"])
        assert np.array_equal(g.edge_norm.cpu().numpy(), gd["norm"].numpy())


# ------------------------------------------------------------------------------ ops vs golden
LP_OPS = ['pre_mult', 'pre_sub', 'pre_add', 'f_zero', 'f_identity', 'f_dense', 'f_dense_comp', 'f_comp',
          'f_sparse', 'f_sparse_comp', 'f_dense_last', 'f_sparse_last', 'a_max', 'a_mean', 'a_sum']


def _run_op(name, dev, g, D, c, state):
    from mr_gnas_b200 import operations_lp as ops
    from mr_gnas_b200.functional import decode_arg
    op = ops.MIXED_OPS[name]({'feature_dim': D, 'drop_aggr': 0.0}).to(dev)
    op.load_state_dict(state)
    x = c["x"].to(dev).requires_grad_(True)
    xin = c["xin"].to(dev).requires_grad_(True)
    out = op(g, x, xin)
    res = {"out": out}
    if out.requires_grad:
        out.backward(c["cot"].to(dev))
    res["dx"], res["dxin"] = x.grad, xin.grad
    res["dparams"] = {k: p.grad for k, p in op.named_parameters()}
    if name == "a_max":
        res["arg"] = decode_arg(g.last_arg)
    return res


@pytest.mark.parametrize("name", LP_OPS)
def test_lp_op_golden(dev, golden_dir, name):
    G = _load(golden_dir, "ops_lp.pt")
    g = _graph_from_golden(G["graph"], dev)
    c = G["cases"][name]
    r = _run_op(name, dev, g, G["D"], c, c["state"])
    _check("out", r["out"], c["out"])
    if c["dx"] is not None:
        _check("dx", r["dx"], c["dx"])
    if c["dxin"] is not None:
        got = r["dxin"] if r["dxin"] is not None else torch.zeros_like(c["dxin"])
        _check("dxin", got, c["dxin"])
    for k, gref in c["dparams"].items():
        if gref is not None:
            _check("d" + k, r["dparams"][k], gref)
    if name == "a_max":
        assert torch.equal(r["arg"].cpu().long(), c["arg"]), "argmax edge ids differ"


# ------------------------------------------------------------------------------ ops vs oracle, larger
@pytest.mark.parametrize("name", ['pre_sub', 'pre_mult', 'f_sparse_comp', 'f_dense_comp', 'f_comp', 'f_sparse_last',
                                  'a_max', 'a_mean', 'a_sum'])
@pytest.mark.parametrize("D", [64, 200, 256])
def test_lp_op_oracle_seeded(dev, name, D):
    from mr_gnas_b200 import operations_lp as ops
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.functional import decode_arg
    N, R, T = 3000, 9, 20000
    trip = O.synth_kg(N, R, T, seed=D)
    ref = O.build_graph(N, trip, R)
    g = MRGraph.from_triples(N, trip, R, device=dev)
    E = 2 * T
    M = E + N
    torch.manual_seed(D + len(name))
    op = ops.MIXED_OPS[name]({'feature_dim': D, 'drop_aggr': 0.0})
    for m in op.modules():
        if isinstance(m, torch.nn.Linear):
            torch.nn.init.xavier_normal_(m.weight)
    rows = N if name.endswith("_last") else M
    x = torch.randn(rows, D)
    if name.startswith("a_"):
        x = torch.relu(x)
    xin = torch.randn(rows, D)
    cot = torch.randn(N if (name.startswith("a_") or name.endswith("_last")) else M, D)
    if name == "a_mean":
        # ReLU'(m) is discontinuous at m = 0: a pre-activation within GEMM rounding distance of zero (the edge-tile
        # GEMM runs as 3xTF32 on the tensor cores, ~1e-6 relative to MKL's fp32) may be gated differently by the
        # two implementations and moves a whole gradient row.  Re-draw the few message rows that have such an
        # entry so that the comparison is about arithmetic, not about which side of zero a rounding error fell.
        W0, b0 = op.linear.weight.detach(), op.linear.bias.detach()
        for _ in range(20):
            m0 = torch.nn.functional.linear(x[:E], W0, b0)
            bad = (m0.abs() < 1e-4 * float(m0.abs().max())).any(1).nonzero().view(-1)
            if bad.numel() == 0:
                break
            x[bad] = torch.relu(torch.randn(bad.numel(), D))
        assert bad.numel() == 0
    # oracle (CPU fp32)
    P = {"op." + k: v.detach().clone().requires_grad_(True) for k, v in op.state_dict().items()}
    xo, xino = x.clone().requires_grad_(True), xin.clone().requires_grad_(True)
    dst = torch.from_numpy(ref["dst"])
    norm = torch.from_numpy(ref["norm"])
    out_o = O.apply_op_lp(name, P, "op", xo, xino, E, norm, dst, N)
    out_o.backward(cot)
    # CUDA
    opg = op.to(dev)
    xg, xing = x.to(dev).requires_grad_(True), xin.to(dev).requires_grad_(True)
    out_g = opg(g, xg, xing)
    out_g.backward(cot.to(dev))
    _check("out", out_g, out_o)
    if name == "a_max":
        # argmax: bit-exact tie-break is checked in test_amax_* / golden; here the CPU and GPU GEMMs round
        # differently, so a near-tie may pick a different edge.  Gradients are therefore validated against
        # the routing the GPU itself reported (exact), and the routing against the oracle's (>= 99.5 %).
        _, arg_o = O.a_op_lp(name, {k: v.detach() for k, v in P.items()}, "op", x, E, dst, N, return_arg=True)
        arg_g = decode_arg(g.last_arg).cpu().long()
        assert (arg_g == arg_o).float().mean() > 0.995
        exp = _amax_expected_grads(x, P["op.linear.weight"].detach(), P["op.linear.bias"].detach(), arg_g, E, cot)
        _check("dx", xg.grad, exp[0])
        _check("dW", opg.linear.weight.grad, exp[1])
        _check("db", opg.linear.bias.grad, exp[2])
        return
    _check("dx", xg.grad, xo.grad)
    if xino.grad is not None:
        _check("dxin", xing.grad if xing.grad is not None else torch.zeros_like(xin), xino.grad)
    for k, p in opg.named_parameters():
        _check("d" + k, p.grad, P["op." + k].grad)


def _amax_expected_grads(x, W, b, arg, E, cot):
    """CPU autograd through out[n,f] = relu(W x_e + b)[arg[n,f], f] + x[E+n, f] for a GIVEN routing arg."""
    xo = x.clone().requires_grad_(True)
    Wo, bo = W.clone().requires_grad_(True), b.clone().requires_grad_(True)
    m = torch.relu(torch.nn.functional.linear(xo[:E], Wo, bo))
    N, D = arg.shape
    cols = torch.arange(D).view(1, -1).expand(N, D)
    picked = torch.where(arg >= 0, m[arg.clamp(min=0), cols], torch.zeros(N, D))
    out = picked + xo[E:]
    out.backward(cot)
    return xo.grad, Wo.grad, bo.grad


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("D", [64, 136, 200, 256])
def test_amax_tc_reads_its_input_through_lazy_bn(dev, D, prec):
    """The fused a_max kernels read x through relu(scale * x + shift) (the lazy BatchNorm of the producing state,
    DESIGN 3); the result must equal the same kernel on the materialised activation (one fmaf per element either
    way, so bit-identical), rows / residual included.  D > 128 runs the CTA-pair kernel, D <= 128 the single-CTA one."""
    from mr_gnas_b200 import functional as K
    from mr_gnas_b200._lib import act, call, ptr, stream
    from mr_gnas_b200.graph import MRGraph
    N, R, T = 3000, 9, 20011
    g = MRGraph.from_triples(N, O.synth_kg(N, R, T, seed=D), R, device=dev)
    E, M = g.E, g.M
    torch.manual_seed(D)
    x = torch.randn(M, D, device=dev)
    a = torch.rand(D, device=dev) + 0.5
    b = torch.randn(D, device=dev) * 0.3
    W = torch.randn(D, D, device=dev) / D ** 0.5
    bias = torch.randn(D, device=dev) * 0.1
    xm = torch.relu(torch.addcmul(b, x, a))          # fused multiply-add on the device, as the kernel's fmaf
    ws = K._tc_workspace(N, D, dev)
    name = "mrg_amax_tc_fwd_bf16" if prec == "bf16" else "mrg_amax_tc_fwd"
    outs = []
    for xa, ra in ((act(x, a, b, True), act(x[E:], a, b, True)), (act(xm), act(xm[E:]))):
        out = torch.empty(N, D, device=dev)
        arg = torch.empty(N, D, dtype=torch.int32, device=dev)
        call(name, xa, ptr(W), ptr(bias), ptr(g.csr.idx), ptr(g.dst), E, N, D, ra, ptr(out), ptr(arg), ptr(ws),
             ws.numel(), stream())
        outs.append((out.clone(), arg.clone()))
    torch.cuda.synchronize()
    assert bool((outs[0][0] == outs[1][0]).all()) and bool((outs[0][1] == outs[1][1]).all())
    ref = torch.relu(xm[:E].double() @ W.double().t() + bias.double())
    full = torch.zeros(N, D, dtype=torch.float64, device=dev).index_reduce_(0, g.dst.long(), ref, "amax", include_self=True)
    full += xm[E:].double()
    tol = 2e-2 if prec == "bf16" else 1e-5
    assert float((outs[0][0].double() - full).abs().max()) <= tol * float(full.abs().max())


@pytest.mark.parametrize("D", [8, 64, 128, 200, 256])
@pytest.mark.parametrize("shape", ["zipf", "hub", "tiny"])
def test_amax_tensor_core_vs_simt_and_oracle(dev, D, shape):
    """Fused tcgen05 a_max (3xTF32) against (a) the oracle in fp32/fp64 and (b) the SIMT/cuBLAS path of
    this library; argmax bit-exact against an exact recomputation on the kernel's own fp32 messages is
    not possible (messages never leave the SM), so arg is checked where the max is unambiguous."""
    from mr_gnas_b200 import operations_lp as ops
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.functional import decode_arg
    if shape == "zipf":
        N, R, T = 3000, 9, 20000
        trip = O.synth_kg(N, R, T, seed=D)
    elif shape == "hub":   # one destination with thousands of in-edges + many isolated nodes
        N, R, T = 2000, 3, 6000
        trip = O.synth_kg(N, R, T, seed=D)
        trip[:4000, 2] = 7
        trip[:, 0] = trip[:, 0] % 500
    else:
        N, R, T = 5, 2, 3
        trip = np.array([[0, 0, 1], [2, 1, 1], [3, 0, 1]])
    ref = O.build_graph(N, trip, R)
    g = MRGraph.from_triples(N, trip, R, device=dev)
    E, M = 2 * len(trip), 2 * len(trip) + N
    torch.manual_seed(D)
    op = ops.a_max_op({'feature_dim': D})
    torch.nn.init.xavier_normal_(op.linear.weight)
    op.linear.bias.data.normal_(0, 0.1)
    x = torch.relu(torch.randn(M, D))
    cot = torch.randn(N, D)
    dst = torch.from_numpy(ref["dst"])
    P = {"op." + k: v.detach().clone().requires_grad_(True) for k, v in op.state_dict().items()}
    xo = x.clone().requires_grad_(True)
    out_o, arg_o = O.a_op_lp("a_max", P, "op", xo, E, dst, N, return_arg=True)
    out_o.backward(cot)
    P64 = {k: v.detach().double() for k, v in P.items()}
    out_64 = O.a_op_lp("a_max", P64, "op", x.double(), E, dst, N)
    opg = op.to(dev)
    res = {}
    for use_tc in (True, False):
        ops.USE_TENSOR_CORES = use_tc
        opg.zero_grad()
        xg = x.to(dev).requires_grad_(True)
        out_g = opg(g, xg, None)
        out_g.backward(cot.to(dev))
        res[use_tc] = (out_g.detach().cpu(), xg.grad.cpu(), opg.linear.weight.grad.cpu().clone(),
                       opg.linear.bias.grad.cpu().clone(), decode_arg(g.last_arg).cpu().long())
    ops.USE_TENSOR_CORES = True
    out_tc, dx_tc, dw_tc, db_tc, arg_tc = res[True]
    out_si, dx_si, dw_si, db_si, arg_si = res[False]
    e_cpu = _err(out_o, out_64.float())
    e_tc = _err(out_tc, out_64.float())
    print(f"D={D} {shape}: fwd err vs fp64: tensor-core {e_tc:.2e}, cpu fp32 {e_cpu:.2e}, simt {_err(out_si, out_64.float()):.2e}")
    _check("out_tc", out_tc, out_o)
    assert e_tc <= max(1e-5, 4 * e_cpu)
    agree = (arg_tc == arg_o).float().mean().item()
    assert agree > 0.995, agree
    # isolated nodes: -1 ; and the argmax must be a real in-edge of that destination
    deg = torch.bincount(dst, minlength=N)
    assert bool((arg_tc[deg == 0] == -1).all())
    nz = arg_tc >= 0
    assert bool((dst[arg_tc[nz]] == torch.arange(N).view(-1, 1).expand(N, D)[nz]).all())
    # gradients: exact against CPU autograd through the routing each path reported
    for tag, (dx_, dw_, db_, arg_) in {"tc": (dx_tc, dw_tc, db_tc, arg_tc), "simt": (dx_si, dw_si, db_si, arg_si)}.items():
        exp = _amax_expected_grads(x, P["op.linear.weight"].detach(), P["op.linear.bias"].detach(), arg_, E, cot)
        _check(tag + " dx", dx_, exp[0])
        _check(tag + " dw", dw_, exp[1])
        _check(tag + " db", db_, exp[2])


def test_bn_act_matches_torch(dev):
    from mr_gnas_b200 import functional as K
    torch.manual_seed(0)
    for rows, D in [(5000, 200), (33, 64), (100000, 128)]:
        bn_ref = torch.nn.BatchNorm1d(D)
        bn_ref.weight.data.uniform_(0.5, 1.5)
        bn_ref.bias.data.uniform_(-0.5, 0.5)
        bn_gpu = torch.nn.BatchNorm1d(D).to(dev)
        bn_gpu.load_state_dict(bn_ref.state_dict())
        y = torch.randn(rows, D) * 2 + 0.7
        cot = torch.randn(rows, D)
        yo = y.clone().double().requires_grad_(True)
        bn64 = torch.nn.BatchNorm1d(D).double()
        bn64.load_state_dict(bn_ref.state_dict())
        so = torch.relu(bn64(yo))
        so.backward(cot.double())
        yg = y.to(dev).requires_grad_(True)
        sg = K.bn_act(yg, bn_gpu, relu=True)
        sg.backward(cot.to(dev))
        _check("s", sg, so.float())
        _check("dy", yg.grad, yo.grad.float())
        _check("dgamma", bn_gpu.weight.grad, bn64.weight.grad.float())
        _check("dbeta", bn_gpu.bias.grad, bn64.bias.grad.float())
        _check("running_mean", bn_gpu.running_mean, bn64.running_mean.float())
        _check("running_var", bn_gpu.running_var, bn64.running_var.float())
        assert int(bn_gpu.num_batches_tracked) == 1


def test_sigmoid_bce(dev):
    from mr_gnas_b200 import functional as K
    torch.manual_seed(1)
    B, N = 64, 1237
    logit = (torch.randn(B, N) * 6)
    logit[0, :5] = torch.tensor([200., -200., 90., -90., 0.])  # exercises the -100 log clamp
    y = (torch.rand(B, N) < 0.01).float() * 0.9 + 1.0 / N
    lo = logit.clone().requires_grad_(True)
    loss_o = O.bce_loss(torch.sigmoid(lo), y)
    loss_o.backward()
    lt = logit.clone().requires_grad_(True)
    loss_t = torch.nn.BCELoss()(torch.sigmoid(lt), y)
    loss_t.backward()
    lg = logit.to(dev).requires_grad_(True)
    loss_g = K.SigmoidBCE.apply(lg, y.to(dev))
    loss_g.backward()
    _check("loss", loss_g.view(1), loss_t.view(1))
    _check("loss_oracle", loss_g.view(1), loss_o.view(1))
    _check("dlogit", lg.grad * (B * N), lt.grad * (B * N))


def test_fail_loudly_on_cpu_tensor():
    from mr_gnas_b200 import functional as K
    with pytest.raises(RuntimeError):
        K.ComposeRows.apply(torch.randn(4, 8), torch.randn(4, 8), 0)


# ------------------------------------------------------------------------------ fused gate backward (TMA row pipeline)
@pytest.mark.parametrize("D", [64, 200, 256])
@pytest.mark.parametrize("variant", ["same_lazy_stats_acc", "split_lazy_stats_acc", "split_plain", "noin_lazy"])
def test_sparse_gate_bwd_fused_vs_fp64(dev, D, variant):
    """mrg_sparse_gate_bwd_fused (lazy BatchNorm backward on load, BN-backward statistics of the input state,
    in-place accumulation) against the same algebra in fp64 torch: operations_lp.py:304-343 backward composed
    with native_batch_norm_backward + threshold_backward on both sides."""
    from mr_gnas_b200 import _lib
    from mr_gnas_b200._lib import act, call, grad, ptr, stream
    lib = _lib.load()
    rows = 1003                       # not a multiple of the 8-row tile
    torch.manual_seed(D + len(variant))
    has_in = not variant.startswith("noin")
    same = variant.startswith("same")
    lazy = "lazy" in variant
    want_stats = "stats" in variant
    acc = 0
    if "acc" in variant:
        acc = 1 if same else 3
    f32 = dict(dtype=torch.float32, device=dev)
    ds, yk, yx, yi = (torch.randn(rows, D, **f32) for _ in range(4))
    if same:
        yi = yx
    ak, bk, ax, bx, ai, bi = (torch.randn(D, **f32) * s + o for s, o in ((0.5, 1), (0.5, 0), (0.5, 1), (0.5, 0), (0.5, 1), (0.5, 0)))
    if same:
        ai, bi = ax, bx
    coef = torch.randn(3 * D, **f32) * 0.3
    gate = torch.rand(rows, **f32)
    rs = torch.rand(rows, **f32)
    v1, v2 = torch.randn(D, **f32), torch.randn(D, **f32)
    dx_old, di_old = torch.randn(rows, D, **f32), torch.randn(rows, D, **f32)
    # ---- fp64 reference
    d = lambda t: t.double()
    dy = d(ds)
    if lazy:
        dz = dy * (d(ak) * d(yk) + d(bk) > 0)
        dy = d(coef[2 * D:]) * dz + d(coef[:D]) + d(coef[D:2 * D]) * d(yk)
    x = torch.relu(d(ax) * d(yx) + d(bx))
    xin = torch.relu(d(ai) * d(yi) + d(bi)) if has_in else None
    sc = d(rs) / 3.0
    dot = (dy * x).sum(1)
    dt = sc * d(gate) * (1 - d(gate)) * dot
    dx_ref = (sc * d(gate)).unsqueeze(1) * dy + dt.unsqueeze(1) * d(v1)
    di_ref = dt.unsqueeze(1) * d(v2) if has_in else None
    if same:
        dx_ref = dx_ref + di_ref
    if acc & 1:
        dx_ref = dx_ref + d(dx_old)
    if has_in and not same and (acc & 2):
        di_ref = di_ref + d(di_old)
    dv1_ref = (dt.unsqueeze(1) * x).sum(0)
    dv2_ref = (dt.unsqueeze(1) * xin).sum(0) if has_in else None
    dzx = dx_ref * (x > 0)
    st_ref = torch.stack([dzx.sum(0), (dzx * d(yx)).sum(0)])
    # ---- kernel
    dx = dx_old.clone() if (acc & 1) else torch.empty(rows, D, **f32)
    if same:
        dxin = dx
    else:
        dxin = (di_old.clone() if (acc & 2) else torch.empty(rows, D, **f32)) if has_in else None
    nparts = int(lib.mrg_stats_nparts(rows))
    dparam = torch.empty(int(lib.mrg_gate_dparam_count(rows, D)), dtype=torch.float64, device=dev)
    xst = torch.empty(nparts * 2 * D, dtype=torch.float64, device=dev) if want_stats else None
    gview = grad(ds, act(yk, ak, bk, True), coef) if lazy else grad(ds)
    xin_view = act(yi, ai, bi, True) if has_in else act(None)
    call("mrg_sparse_gate_bwd_fused", gview, act(yx, ax, bx, True), xin_view, ptr(gate), rows, D, ptr(v1),
         ptr(v2) if has_in else None, ptr(rs), 1.0 / 3.0, ptr(dx), ptr(dxin) if has_in else None, acc, ptr(dparam),
         ptr(xst), stream())
    torch.cuda.synchronize()
    _check("dx", dx, dx_ref.float())
    if has_in and not same:
        _check("dxin", dxin, di_ref.float())
    dp = dparam.view(nparts, 2 * D + 1).sum(0)
    _check("dv1", dp[:D].float(), dv1_ref.float())
    if has_in:
        _check("dv2", dp[D:2 * D].float(), dv2_ref.float())
    _check("dc", dp[2 * D:].float(), dt.sum().view(1).float())
    if want_stats:
        _check("x bwd stats", xst.view(nparts, 2, D).sum(0).float(), st_ref.float())


# ------------------------------------------------------------------------------ fused DistMult + BCE (tcgen05)
@pytest.mark.parametrize("B,N,D", [(256, 14541, 200), (32, 700, 64), (300, 1000, 128), (1, 5, 8), (130, 257, 256)])
def test_distmult_bce_fused_vs_fp64(dev, B, N, D):
    """mrg_distmult_bce_fwd (3xTF32 tcgen05 GEMM + sigmoid + BCE epilogue) and its backward against fp64 torch:
    sf_DisMult_op.forward (operations_lp.py:115-127) + nn.BCELoss; ragged last tile, B > 256 (two launches),
    B < 128 (one accumulator half)."""
    from mr_gnas_b200 import functional as K
    torch.manual_seed(B + N + D)
    ent = (torch.randn(N, D, device=dev) * 0.3).requires_grad_(True)
    sub = (torch.randn(B, D, device=dev) * 0.5).requires_grad_(True)
    rel = (torch.randn(B, D, device=dev) * 0.5).requires_grad_(True)
    label = ((torch.rand(B, N, device=dev) < 0.05).float() * 0.9 + 1.0 / N).clamp(max=1.0)
    loss = K.DistMultBCE.apply(ent, sub, rel, label)
    loss.backward()
    e64, s64, r64 = (t.detach().double().requires_grad_(True) for t in (ent, sub, rel))
    logit64 = (s64 * r64) @ e64.t()
    ref = torch.nn.functional.binary_cross_entropy(torch.sigmoid(logit64), label.double())
    ref.backward()
    _check("loss", loss.view(1), ref.view(1).float())
    _check("d all_ent", ent.grad, e64.grad.float())
    _check("d sub_emb", sub.grad, s64.grad.float())
    _check("d rel_emb", rel.grad, r64.grad.float())
    # and against the unfused library path of this repo on the same inputs
    loss_u = K.SigmoidBCE.apply(torch.mm((sub * rel).detach(), ent.detach().t()), label)
    _check("loss vs cuBLAS + sigmoid_bce kernel", loss.view(1), loss_u.view(1))


# ------------------------------------------------------------------------------ node-level Linear (tcgen05, 3xTF32)
@pytest.mark.parametrize("rows,K,F", [(14541, 800, 200), (14541, 200, 200), (3000, 64, 64), (1500, 256, 520), (2000, 40, 8)])
def test_linear_tc_vs_fp64(dev, rows, K, F):
    """mrg_linear_tc_fwd (forward and, with W^T, the input gradient) against fp64: nn.Linear of the cell's `concat`
    (model_lp.py:70-71) and `linear_e` (:124); F > 256 runs several launches, K not a multiple of 32 is zero padded (K % 8 == 0 is required)."""
    from mr_gnas_b200 import functional as K_
    torch.manual_seed(rows + K + F)
    lin = torch.nn.Linear(K, F).to(dev)
    x = torch.randn(rows, K, device=dev, requires_grad=True)
    y = K_.LinearTC.apply(x, lin.weight, lin.bias)
    cot = torch.randn(rows, F, device=dev)
    y.backward(cot)
    x64 = x.detach().double().requires_grad_(True)
    w64, b64 = lin.weight.detach().double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    y64 = x64 @ w64.t() + b64
    y64.backward(cot.double())
    _check("y", y, y64.float())
    _check("dx", x.grad, x64.grad.float())
    _check("dw", lin.weight.grad, w64.grad.float())
    _check("db", lin.bias.grad, b64.grad.float())


@pytest.mark.parametrize("rows,F1,F2,kmajor", [
    (14541, 200, 800, False), (14541, 200, 200, False),      # weight gradients of `concat` and `linear_e` at C1
    (14541, 256, 200, True),                                  # DistMult dq = dl . ent   (dl [B, N] is the K-major operand)
    (256, 14541, 200, False),                                 # DistMult dent = dl^T . q (odd leading dimension: scalar loads)
    (16, 803, 64, False), (1000, 64, 64, False), (2500, 520, 136, False), (33, 8, 8, True), (4097, 24, 300, True)])
def test_gemm_red_vs_fp64(dev, rows, F1, F2, kmajor):
    """mrg_gemm_red: C = A^T B with the reduction over the rows (3xTF32, split over the rows, deterministic fold)
    against fp64, the fp32 library GEMM as the yardstick; ragged tiles, rows not a multiple of 32, odd leading
    dimensions, both operand forms; bit-identical when repeated."""
    from mr_gnas_b200._lib import call, ptr, stream, load
    torch.manual_seed(rows + F1 + F2)
    A = torch.randn(rows, F1, device=dev)
    B = torch.randn(rows, F2, device=dev)
    Ain = A.t().contiguous() if kmajor else A
    bias = torch.randn(F2, device=dev)
    lib = load()
    ws = torch.empty(max(int(lib.mrg_gemm_red_workspace_bytes(rows, F1, F2)), 16), dtype=torch.uint8, device=dev)
    outs = []
    for _ in range(2):
        C = torch.full((F1, F2 + 3), 7.0, device=dev)          # ldc > F2: the padding columns must stay untouched
        cs = torch.full((F1,), 7.0, device=dev)
        call("mrg_gemm_red", ptr(Ain), Ain.shape[1], 1 if kmajor else 0, ptr(B), F2, rows, F1, F2, ptr(C), F2 + 3,
             ptr(cs), ptr(bias), ptr(ws), ws.numel(), stream())
        outs.append(C)
        sums = cs
    torch.cuda.synchronize()
    assert bool((outs[0] == outs[1]).all())
    assert bool((outs[0][:, F2:] == 7.0).all())
    ref = A.double().t() @ B.double() + bias.double()
    lib32 = A.t() @ B + bias
    scale = float(ref.abs().max())
    e_tc = float((outs[0][:, :F2].double() - ref).abs().max()) / scale
    e_lib = float((lib32.double() - ref).abs().max()) / scale
    print(f"gemm_red rows={rows} F1={F1} F2={F2} kmajor={kmajor}: err vs fp64 {e_tc:.2e} (library fp32 GEMM {e_lib:.2e})")
    assert e_tc <= max(1e-5, 4 * e_lib)
    cref = A.double().sum(0)         # the virtual ones column: column sums of A (bias gradient)
    assert float((sums.double() - cref).abs().max()) <= 1e-5 * max(float(cref.abs().max()), float(A.abs().sum(0).max()) * 1e-2)


def test_matmul_tc_vs_fp64(dev):
    """K.matmul (rel_wt @ embedding_e.weight and rel_embed @ w_rel, model_lp.py:125,133) forward and both gradients
    on mrg_gemm_red against fp64."""
    from mr_gnas_b200 import functional as K_
    torch.manual_seed(9)
    K_.USE_TC_MATMUL = True          # off by default (see functional.py): the library GEMM serves these tiny products
    for m, k, n in ((475, 475, 200), (475, 200, 200), (23, 23, 64)):
        X = torch.randn(m, k, device=dev, requires_grad=True)
        Y = torch.randn(k, n, device=dev, requires_grad=True)
        cot = torch.randn(m, n, device=dev)
        C = K_.matmul(X, Y)
        C.backward(cot)
        X64, Y64 = X.detach().double().requires_grad_(True), Y.detach().double().requires_grad_(True)
        C64 = X64 @ Y64
        C64.backward(cot.double())
        _check("C", C, C64.float())
        _check("dX", X.grad, X64.grad.float())
        _check("dY", Y.grad, Y64.grad.float())
    K_.USE_TC_MATMUL = False


def test_gemm_red_zero_rows(dev):
    from mr_gnas_b200._lib import call, ptr, stream
    C = torch.full((8, 16), 3.0, device=dev)
    A = torch.zeros(1, 8, device=dev)
    B = torch.zeros(1, 16, device=dev)
    ws = torch.empty(16, dtype=torch.uint8, device=dev)
    call("mrg_gemm_red", ptr(A), 8, 0, ptr(B), 16, 0, 8, 16, ptr(C), 16, None, None, ptr(ws), ws.numel(), stream())
    torch.cuda.synchronize()
    assert bool((C == 0).all())


@pytest.mark.parametrize("D", [64, 200, 256])
def test_amax_bf16_variant(dev, D):
    """Reduced-precision variant of the fused a_max forward (mrg_amax_tc_fwd_bf16: activated rows and W rounded to
    bf16, fp32 accumulation) against the fp32-class 3xTF32 kernel on the same inputs.
    STATED TOLERANCE (bf16 class): max|out_bf16 - out_fp32| <= 2e-2 * max|out_fp32| and 2-norm error <= 1e-2; the
    routing (argmax edge per destination and feature) must agree on >= 90 % of the pairs (a bf16-rounded message
    may pick another near-maximal edge); isolated destinations and ReLU-gated maxima are encoded identically.
    The backward is the fp32 sparse kernel fed with the routing the forward reported, so gradients are exact for
    that routing (checked against CPU autograd through it)."""
    from mr_gnas_b200 import functional as K
    from mr_gnas_b200 import operations_lp as ops
    from mr_gnas_b200.graph import MRGraph
    from parity import norm_err, rel_err
    N, R, T = 3000, 9, 20000
    trip = O.synth_kg(N, R, T, seed=D)
    g = MRGraph.from_triples(N, trip, R, device=dev)
    E = 2 * T
    torch.manual_seed(D)
    op = ops.MIXED_OPS['a_max']({'feature_dim': D, 'drop_aggr': 0.0}).to(dev)
    x = torch.relu(torch.randn(E + N, D, device=dev))
    cot = torch.randn(N, D, device=dev)
    res = {}
    for prec in ("fp32", "bf16"):
        K.AMAX_PRECISION = prec
        try:
            xg = x.clone().requires_grad_(True)
            op.zero_grad()
            out = op(g, xg, xg)
            out.backward(cot)
            res[prec] = (out.detach(), K.decode_arg(g.last_arg).clone(), g.last_arg.clone(), xg.grad.clone(),
                         op.linear.weight.grad.clone(), op.linear.bias.grad.clone())
        finally:
            K.AMAX_PRECISION = "fp32"
    o32, a32, enc32 = res["fp32"][:3]
    o16, a16, enc16 = res["bf16"][:3]
    e_max, e_nrm = rel_err(o16, o32), norm_err(o16, o32)
    agree = float((a16 == a32).float().mean())
    print(f"bf16 a_max D={D}: max-norm rel err {e_max:.2e}, 2-norm rel err {e_nrm:.2e}, routing agreement {agree:.4f}")
    assert e_max <= 2e-2 and e_nrm <= 1e-2
    assert agree >= 0.90
    assert torch.equal(enc16 == -1, enc32 == -1)                      # isolated destinations
    # gradients are exact for the routing AND the ReLU gating the bf16 forward reported (enc >= 0: the routed
    # message was positive in bf16 arithmetic): CPU autograd through out[n,f] = (W x_e + b)[enc[n,f], f] + x[E+n, f]
    xo = x.cpu().clone().requires_grad_(True)
    Wo = op.linear.weight.detach().cpu().clone().requires_grad_(True)
    bo = op.linear.bias.detach().cpu().clone().requires_grad_(True)
    m = torch.nn.functional.linear(xo[:E], Wo, bo)
    enc = enc16.cpu().long()
    cols = torch.arange(D).view(1, -1).expand(N, D)
    picked = torch.where(enc >= 0, m[enc.clamp(min=0), cols], torch.zeros(N, D))
    (picked + xo[E:]).backward(cot.cpu())
    _check("dx", res["bf16"][3], xo.grad)
    _check("dW", res["bf16"][4], Wo.grad)
    _check("db", res["bf16"][5], bo.grad)
