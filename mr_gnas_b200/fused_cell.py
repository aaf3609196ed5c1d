"""Fused executor for the edge-level part of a genotype cell (reference dataflow: model_lp.py:59-74,
SURVEY.md appendix B).

The module-by-module path materialises, per op, the op output, its BatchNorm output and its ReLU output
as [E+N, D] tensors and lets autograd add full-size gradient tensors.  Here the whole chain
    gather -> pre_* -> BN/ReLU -> f_sparse_comp ... -> a_max | a_sum
is one autograd node that
  * stores only the pre-BN output y_k of each op; BN+ReLU is applied on load by every consumer
    (`mrg_act`), so each edge-level op costs one read per input and one write;
  * never materialises the gathered inputs (fused into the compose kernel) nor the [E, D] messages
    (fused tcgen05 a_max);
  * runs the backward with explicit, in-place accumulated gradient buffers, the sparse a_max backward
    and deterministic segmented reductions for the gather.
Supported pattern: node 1 = pre_* of the gather; every further edge-level node has exactly one producer,
which is f_sparse_comp(x = earlier edge state, x_in = node 1); aggregators are a_max / a_sum of an
edge-level state.  Anything else falls back to the per-module path (still CUDA, just less fused)."""
import torch

from . import _lib
from . import dist as D_
from . import functional as K
from ._lib import act, call, ptr, stream


def plan_for(cell):
    """-> (pre OpModule, [(node, gate OpModule, src node)], [(node, agg OpModule, src node)]) or None."""
    ops = cell._ops
    nb = cell._nb_nodes
    first = ops[0][0][0] if len(ops[0][0]) else None
    if first is None or not hasattr(first.op, 'comp') or cell.uses_state0():
        return None
    level = {1: 'edge'}
    gates, aggs = [], []
    for n in range(1, nb):
        prods = [(i, ops[n][i][0]) for i in range(n + 1) if len(ops[n][i]) > 0]
        node = n + 1
        if len(prods) != 1:
            if any(level.get(i) == 'edge' for i, _ in prods):
                return None
            level[node] = 'node'
            continue
        i, om = prods[0]
        if level.get(i) == 'edge':
            if om.op_name == 'f_sparse_comp':
                gates.append((node, om, i))
                level[node] = 'edge'
            elif om.op_name in ('a_max', 'a_sum') and not (om.op_name == 'a_sum' and om.op.training and om.op.drop_aggr > 0):
                if om.op_name == 'a_max' and not K.amax_tc_supported(cell._feature_dim):
                    return None
                aggs.append((node, om, i))
                level[node] = 'node'
            else:
                return None
        else:
            level[node] = 'node'
    if any(level.get(c) == 'edge' for c in cell._concat_node) or not aggs:
        return None
    return first, gates, aggs


class EdgeChain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, plan, training, ent, rel, *params):
        first, gates, aggs = plan
        ent, rel = K._f32c(ent), K._f32c(rel)
        K.check_tables(g, ent, rel)
        D = ent.shape[1]
        dev = ent.device
        M, E, N = g.M, g.E, g.N
        it = iter(params)
        lib = _lib.load()

        def bn_affine(om, stats, rows):
            """training: finalize batch stats (updates running stats); eval: running stats."""
            bn = om.batchnorm_h
            gamma, beta = next(it), next(it)
            if om.op_name == 'pre_mult':       # model_lp.py:31: no BN/ReLU after pre_mult
                return None, gamma, beta
            if training:
                bn.num_batches_tracked.add_(1)
                a = torch.empty(D, dtype=torch.float32, device=dev)
                b, mean, invstd = torch.empty_like(a), torch.empty_like(a), torch.empty_like(a)
                mom = K.bn_momentum(bn)
                stats, nparts_, rows = D_.sync_stats(g.part, stats, stats.numel() // (2 * D), 2 * D, rows)
                call("mrg_bn_finalize", ptr(stats), nparts_, rows, D, ptr(gamma), ptr(beta), float(bn.eps),
                     float(mom), ptr(bn.running_mean), ptr(bn.running_var), ptr(mean), ptr(invstd), ptr(a), ptr(b),
                     stream())
            else:
                invstd = torch.rsqrt(bn.running_var + bn.eps)
                mean = bn.running_mean
                a = (gamma * invstd).contiguous()
                b = (beta - a * mean).contiguous()
            return (a, b, mean, invstd), gamma, beta

        Y, BN, GATE = {}, {}, {}
        # node 1: gather + compose (+ column statistics)
        y1 = torch.empty(M, D, dtype=torch.float32, device=dev)
        st = K._stats_buf(K.stats_nparts(M), D, dev)
        b = 4 * D
        call("mrg_compose_fwd", ptr(ent), ptr(g.src_final), ptr(rel), ptr(g.et_final), M, D, first.op.comp, ptr(y1),
             ptr(st), stream(), nbytes=M * (b + 8))
        Y[1] = y1
        BN[1] = bn_affine(first, st, M)

        def view(k, lo=0, hi=None):
            aff = BN[k][0]
            y = Y[k] if (lo == 0 and hi is None) else Y[k][lo:hi]
            return act(y) if aff is None else act(y, aff[0], aff[1], True)

        bounds = [(0, g.half), (g.half, E), (E, M)]
        norm = g.norm()
        gate_params = {}
        for node, om, i in gates:
            v1, v2, c = K._f32c(next(it)), K._f32c(next(it)), K._f32c(next(it))
            gate_params[node] = (v1, v2, c)
            y = torch.empty(M, D, dtype=torch.float32, device=dev)
            gt = torch.empty(M, dtype=torch.float32, device=dev)
            nparts = [K.stats_nparts(hi - lo) for lo, hi in bounds]
            st = K._stats_buf(sum(nparts), D, dev)
            off = 0
            for s_, (lo, hi) in enumerate(bounds):
                rs = norm[lo:] if lo < E else None
                call("mrg_sparse_gate_fwd", view(i, lo, hi), view(1, lo, hi), hi - lo, D, ptr(v1[s_]), ptr(v2[s_]),
                     ptr(c[s_:s_ + 1]), ptr(rs), 1.0 / 3.0, ptr(y[lo:hi]), ptr(gt[lo:hi]), ptr(st[off * 2 * D:]), stream(),
                     nbytes=(hi - lo) * ((2 if i == 1 else 3) * b + 8))
                off += nparts[s_]
            Y[node], GATE[node] = y, gt
            BN[node] = bn_affine(om, st, M)
        outs, ARG, agg_params = [], {}, {}
        for node, om, i in aggs:
            out = torch.empty(N, D, dtype=torch.float32, device=dev)
            if om.op_name == 'a_max':
                W, bias = K._f32c(next(it)), K._f32c(next(it))
                agg_params[node] = (W, bias)
                arg = torch.empty(N, D, dtype=torch.int32, device=dev)
                ws = K._tc_workspace(N, D, dev)
                call(K.amax_tc_call(), view(i), ptr(W), ptr(bias), ptr(g.csr.idx), ptr(g.dst), E, N, D, view(i, E, M),
                     ptr(out), ptr(arg), ptr(ws), ws.numel(), stream(), nbytes=E * (b + 8) + 3 * N * b)
                ARG[node] = arg
                g.last_arg = arg
                if getattr(g, 'arg_trace', None) is not None:   # tests: every a_max's encoded argmax, in order
                    g.arg_trace.append(arg)
            else:
                K.seg_reduce_raw(g.csr, 0, view(i), D, out, residual=view(i, E, M))
            outs.append(out)
        ctx.g, ctx.plan, ctx.training = g, plan, training
        ctx.state = (Y, BN, GATE, ARG, gate_params, agg_params, ent, rel)
        ctx.n_params = len(params)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *douts):
        g, (first, gates, aggs), training = ctx.g, ctx.plan, ctx.training
        Y, BN, GATE, ARG, gate_params, agg_params, ent, rel = ctx.state
        ctx.state = None
        D = ent.shape[1]
        dev = ent.device
        M, E, N = g.M, g.E, g.N
        DS = {}                 # grad w.r.t. the ACTIVATED state k, [M, D], accumulated in place
        grads = {}              # parameter grads keyed like the forward's parameter order

        def view(k, lo=0, hi=None):
            aff = BN[k][0]
            y = Y[k] if (lo == 0 and hi is None) else Y[k][lo:hi]
            return act(y) if aff is None else act(y, aff[0], aff[1], True)

        # ---- aggregators (they overwrite: run them first on each state)
        for (node, om, i), dout in zip(aggs, douts):
            dout = K._f32c(dout)
            fresh = i not in DS
            if om.op_name == 'a_max':
                W, _ = agg_params[node]
                buf = torch.empty(M, D, dtype=torch.float32, device=dev)
                _, dw, db = K.amax_backward(g, dout, ARG[node], view(i), W, M, True, True, dx=buf)
                grads[('agg', node)] = (dw, db)
                if fresh:
                    DS[i] = buf
                else:
                    DS[i].add_(buf)
            else:
                if fresh:
                    DS[i] = torch.empty(M, D, dtype=torch.float32, device=dev)
                call("mrg_seg_reduce_bwd", 0, ptr(dout), None, None, act(None), ptr(g.dst), ptr(g.csr.ptr), E, N, D,
                     ptr(DS[i]), 0 if fresh else 1, stream())
        bounds = [(0, g.half), (g.half, E), (E, M)]
        norm = g.norm()
        lib = _lib.load()
        fused_bwd = bool(lib.mrg_sparse_gate_bwd_fused_supported(D))
        # pending gradient contributions per edge-level state: when the LAST one is the dx of a fused gate
        # backward, that kernel also emits the state's BN-backward column sums (BST) and no reduce pass runs
        pending = {}
        for node, om, i in aggs:
            pending[i] = pending.get(i, 0) + 1
        for node, om, i in gates:
            pending[i] = pending.get(i, 0) + 1
            if i != 1:
                pending[1] = pending.get(1, 0) + 1
        for node, om, i in aggs:
            pending[i] -= 1
        BST = {}

        def bn_coef(k):
            """BN backward statistics of state k -> (dgamma, dbeta), coef [3,D] (None: no BN after this op)."""
            aff, gamma, beta = BN[k]
            if aff is None:
                return (torch.zeros_like(gamma), torch.zeros_like(beta)), None
            a, b, mean, invstd = aff
            if k in BST:
                bst, nparts = BST.pop(k)
            else:
                nparts = K.stats_nparts(M)
                bst = K._stats_buf(nparts, D, dev)
                call("mrg_bn_bwd_reduce", ptr(DS[k]), act(Y[k], a, b, True), M, D, ptr(bst), stream(),
                     nbytes=2 * M * 4 * D)
            dgamma = torch.empty(D, dtype=torch.float32, device=dev)
            dbeta = torch.empty_like(dgamma)
            coef = torch.empty(3 * D, dtype=torch.float32, device=dev)
            bst, nparts, rows_g = D_.sync_stats(g.part, bst, nparts, 2 * D, M)
            call("mrg_bn_bwd_finalize", ptr(bst), nparts, rows_g, D, ptr(gamma), ptr(mean), ptr(invstd), ptr(dgamma),
                 ptr(dbeta), ptr(coef), stream())
            if not training:
                coef = torch.cat([torch.zeros(2 * D, device=dev), a]).contiguous()
            return D_.unshare_param_grads(g.part, dgamma, dbeta), coef

        def bn_apply(k, coef):
            """ds_k -> dy_k in place (only where no fused consumer reads the gradient lazily)."""
            if coef is None:
                return
            aff = BN[k][0]
            call("mrg_bn_bwd_apply", ptr(DS[k]), act(Y[k], aff[0], aff[1], True), ptr(coef), M, D, ptr(DS[k]), 0,
                 stream(), nbytes=3 * M * 4 * D)

        # ---- gates in reverse order
        for node, om, i in reversed(gates):
            if node not in DS:       # state never consumed (dead branch): zero gradient
                DS[node] = torch.zeros(M, D, dtype=torch.float32, device=dev)
            grads[('bn', node)], coef = bn_coef(node)
            if not fused_bwd:
                bn_apply(node, coef)
            dy = DS.pop(node)
            v1, v2, c = gate_params[node]
            same = i == 1
            fresh_x, fresh_in = i not in DS, 1 not in DS
            if fresh_x:
                DS[i] = torch.empty(M, D, dtype=torch.float32, device=dev)
            if fresh_in and not same:
                DS[1] = torch.empty(M, D, dtype=torch.float32, device=dev)
            accum = (0 if fresh_x else 1) | (0 if (fresh_in or same) else 2)
            dv1, dv2 = torch.empty_like(v1), torch.empty_like(v2)
            dc = torch.empty(3, dtype=torch.float32, device=dev)
            last_for_x = fused_bwd and pending[i] == 1 and BN[i][0] is not None
            if last_for_x:
                nparts = [K.stats_nparts(hi - lo) for lo, hi in bounds]
                xst = K._stats_buf(sum(nparts), D, dev)
                BST[i] = (xst, sum(nparts))
            off = 0
            aff_k = BN[node][0]
            for s_, (lo, hi) in enumerate(bounds):
                n = hi - lo
                rs = norm[lo:] if lo < E else None
                dparam = torch.empty(int(lib.mrg_gate_dparam_count(n, D)), dtype=torch.float64, device=dev)
                n_in = (2 if same else 3) + bin(accum).count("1")
                if fused_bwd:
                    yk = act(Y[node][lo:hi], aff_k[0], aff_k[1], True) if coef is not None else None
                    call("mrg_sparse_gate_bwd_fused", _lib.grad(dy[lo:hi], yk, coef), view(i, lo, hi), view(1, lo, hi),
                         ptr(GATE[node][lo:hi]), n, D, ptr(v1[s_]), ptr(v2[s_]), ptr(rs), 1.0 / 3.0, ptr(DS[i][lo:hi]),
                         ptr(DS[1][lo:hi]), accum, ptr(dparam), ptr(xst[off * 2 * D:]) if last_for_x else None, stream(),
                         nbytes=n * (4 * D * (n_in + (1 if coef is not None else 0) + (1 if same else 2)) + 8))
                    if last_for_x:
                        off += nparts[s_]
                else:
                    call("mrg_sparse_gate_bwd", ptr(dy[lo:hi]), view(i, lo, hi), view(1, lo, hi), ptr(GATE[node][lo:hi]),
                         n, D, ptr(v1[s_]), ptr(v2[s_]), ptr(rs), 1.0 / 3.0, ptr(DS[i][lo:hi]), ptr(DS[1][lo:hi]),
                         accum, ptr(dparam), stream(), nbytes=n * (4 * D * (n_in + (1 if same else 2)) + 8))
                call("mrg_sparse_gate_bwd_finalize", ptr(dparam), n, D, ptr(dv1[s_]), ptr(dv2[s_]), ptr(dc[s_:s_ + 1]),
                     stream())
            pending[i] -= 1
            if not same:
                pending[1] -= 1
            grads[('gate', node)] = (dv1, dv2, dc)
            del dy
        # ---- node 1: BN backward, then the gather/compose backward
        if 1 not in DS:
            DS[1] = torch.zeros(M, D, dtype=torch.float32, device=dev)
        grads[('bn', 1)], coef1 = bn_coef(1)
        bn_apply(1, coef1)
        dy1 = DS.pop(1)
        comp = first.op.comp
        dent = drel = None
        if ctx.needs_input_grad[3]:
            dent = torch.empty_like(ent)
            if comp == 1:
                K.seg_reduce_raw(g.csc, 0, act(dy1), D, dent, mul=rel, mul_idx=g.et_final)
            else:
                K.seg_reduce_raw(g.csc, 0, act(dy1), D, dent)
        if ctx.needs_input_grad[4]:
            drel = torch.empty_like(rel)
            if comp == 1:
                K.seg_reduce_raw(g.rel, 0, act(dy1), D, drel, mul=ent, mul_idx=g.src_final)
            else:
                K.seg_reduce_raw(g.rel, 0, act(dy1), D, drel, alpha=-1.0 if comp == 0 else 1.0)
        # ---- parameter grads in forward order
        # forward consumed: bn(1) ; per gate: v1,v2,c then bn ; per a_max: W,b
        ordered = [*grads[('bn', 1)]]
        for node, om, i in gates:
            ordered += list(grads[('gate', node)]) + list(grads[('bn', node)])
        for node, om, i in aggs:
            if om.op_name == 'a_max':
                ordered += list(grads[('agg', node)])
        return (None, None, None, dent, drel, *ordered)


def run(cell, g, ent, rel, plan):
    """Edge-level chain of `cell` -> {node: pre-BN aggregated [N, D] tensor}."""
    first, gates, aggs = plan
    params = [first.batchnorm_h.weight, first.batchnorm_h.bias]
    D = cell._feature_dim
    for node, om, i in gates:
        op = om.op
        v1, v2, c = K.collapse_gates(D, ((op.W_in, op.a_in), (op.W_out, op.a_out), (op.W_self, op.a_self)))
        params += [v1, v2, c, om.batchnorm_h.weight, om.batchnorm_h.bias]
    for node, om, i in aggs:
        if om.op_name == 'a_max':
            params += [om.op.linear.weight, om.op.linear.bias]
    training = first.batchnorm_h.training
    outs = EdgeChain.apply(g, plan, training, ent, rel, *params)
    return {node: o for (node, om, i), o in zip(aggs, outs)}
