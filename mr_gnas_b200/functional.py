"""torch.autograd glue over the libmrgnas C ABI: one Function per kernel family.  torch is
used for device memory, streams and the tape only; all arithmetic on [rows, D] feature
matrices happens in the hand-written CUDA kernels (no eager / CPU fallback)."""
import ctypes
import os

import torch

from . import _lib
from . import dist as _dist
from ._lib import act, call, ptr, stream


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _stats_buf(nparts, D, device):
    return torch.empty(max(nparts, 1) * 2 * D, dtype=torch.float64, device=device)


def stats_nparts(rows):
    return int(_lib.load().mrg_stats_nparts(int(rows)))


# ------------------------------------------------------------------------------------------
# K1: compose (elementwise op form) and gather+compose (fused network form)
# ------------------------------------------------------------------------------------------
class ComposeRows(torch.autograd.Function):
    """pre_sub / pre_mult / pre_add on already gathered rows (operations_lp.py:71-98)."""

    @staticmethod
    def forward(ctx, x, hr, comp):
        x, hr = _f32c(x), _f32c(hr)
        rows, D = x.shape
        y = torch.empty_like(x)
        call("mrg_compose_fwd", ptr(x), None, ptr(hr), None, rows, D, comp, ptr(y), None, stream())
        ctx.comp = comp
        ctx.save_for_backward(*((x, hr) if comp == 1 else ()))
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _f32c(dy)
        rows, D = dy.shape
        x, hr = ctx.saved_tensors if ctx.comp == 1 else (None, None)
        dx = torch.empty_like(dy) if ctx.needs_input_grad[0] else None
        dr = torch.empty_like(dy) if ctx.needs_input_grad[1] else None
        call("mrg_compose_bwd_rows", ptr(dy), ptr(x), ptr(hr), rows, D, ctx.comp, ptr(dx), ptr(dr), stream())
        return dx, dr, None


def seg_reduce_raw(seg, kind, m_act, D, out, arg=None, mul=None, mul_idx=None, alpha=1.0, residual=None,
                   accumulate=False):
    ws = seg.workspace(D, kind)
    res = residual if residual is not None else act(None)
    call("mrg_seg_reduce_fwd", kind, m_act, ptr(seg.ptr), ptr(seg.idx), ptr(seg.chunk_first), ptr(seg.chunk_seg),
         seg.nseg, seg.max_chunks, D, ptr(mul), ptr(mul_idx), float(alpha), res, 1 if accumulate else 0, ptr(out),
         ptr(arg), ptr(ws), ws.numel(), stream(), nbytes=seg.total * (4 * D + 4) + seg.nseg * 4 * D)
    return out


def check_tables(g, ent, rel):
    """The gather reads ent[src_final] / rel[et_final] and its backward writes one row per CSC / relation
    segment: both tables must have exactly the rows the graph was segmented for (else rows of the gradient
    would stay unwritten)."""
    if ent.shape[0] != g.n_src or rel.shape[0] != g.n_rel_rows:
        raise RuntimeError(f"gather tables [{ent.shape[0]}, {rel.shape[0]}] rows do not match the graph "
                           f"(entities {g.n_src}, relation rows {g.n_rel_rows}); see MRGraph.require_tables")


class GatherCompose(torch.autograd.Function):
    """y[i] = h[src_final[i]] (-|*|+) r[et_final[i]] over the M edge-expanded rows, plus BN
    column statistics -- model_lp.py:126-131 fused with operations_lp.py:71-98.  Backward is
    a deterministic segmented sum over the src-CSC (dh) and the relation segments (dr)
    instead of the reference's atomic index_add_."""

    @staticmethod
    def forward(ctx, h, r, g, comp):
        h, r = _f32c(h), _f32c(r)
        D = h.shape[1]
        check_tables(g, h, r)
        y = torch.empty(g.M, D, dtype=torch.float32, device=h.device)
        nparts = stats_nparts(g.M)
        stats = _stats_buf(nparts, D, h.device)
        call("mrg_compose_fwd", ptr(h), ptr(g.src_final), ptr(r), ptr(g.et_final), g.M, D, comp, ptr(y), ptr(stats),
             stream())
        ctx.g, ctx.comp = g, comp
        ctx.save_for_backward(h, r)
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _):
        g, comp = ctx.g, ctx.comp
        h, r = ctx.saved_tensors
        dy = _f32c(dy)
        D = dy.shape[1]
        dh = dr = None
        if ctx.needs_input_grad[0]:
            dh = torch.empty_like(h)
            if comp == 1:
                seg_reduce_raw(g.csc, 0, act(dy), D, dh, mul=r, mul_idx=g.et_final)
            else:
                seg_reduce_raw(g.csc, 0, act(dy), D, dh)
        if ctx.needs_input_grad[1]:
            dr = torch.empty_like(r)
            if comp == 1:
                seg_reduce_raw(g.rel, 0, act(dy), D, dr, mul=h, mul_idx=g.src_final)
            else:
                seg_reduce_raw(g.rel, 0, act(dy), D, dr, alpha=-1.0 if comp == 0 else 1.0)
        return dh, dr, None, None


class GatherRows(torch.autograd.Function):
    """table[idx] over the M edge-expanded rows of a graph whose segment list `seg` groups those rows by idx (the src-CSC
    for entity rows, the relation segments for relation rows): model_search_lp.py:139-145 `all_ent_emb[src_id_final]`,
    `ent_emb[src_in]`, `rel_embed[edge_type_final]`.  The backward is the deterministic segmented sum the LP network's
    fused gather already uses (mrg_seg_reduce_fwd) instead of ATen's sort-based index_put: with ~23 relation rows
    behind 60,000 gathered rows that kernel serialises on the duplicates -- 7 launches of 16.9 ms were 78 % of the GPU
    time of the C3 supernet step at graph_batch_size 30,000 (profiles/r02b_search_c3.jsonl)."""

    @staticmethod
    def forward(ctx, table, idx, seg):
        table = _f32c(table)
        if table.shape[0] != seg.nseg or idx.numel() != seg.total:
            raise RuntimeError(f"gather_rows: table has {table.shape[0]} rows / {idx.numel()} gathered rows, the graph's "
                               f"segment list covers {seg.nseg} / {seg.total}")
        ctx.seg = seg
        return table.index_select(0, idx)

    @staticmethod
    def backward(ctx, dy):
        seg = ctx.seg
        dy = _f32c(dy)
        D = dy.shape[1]
        dt = torch.empty(seg.nseg, D, dtype=torch.float32, device=dy.device)
        seg_reduce_raw(seg, 0, act(dy), D, dt)
        return dt, None, None


def gather_rows(table, idx, seg):
    return GatherRows.apply(table, idx, seg)


def key_segments(keys, nseg, owner, name):
    """Segment list (graph._Segments) grouping the positions of `keys` [rows] by key value in [0, nseg): built with one
    stable sort the first time `owner` (a graph / block object) is seen with this key tensor, then cached on it -- the NC
    path's full-graph blocks are static, so the relation gather's backward costs one segmented sum per step."""
    from .graph import _Segments
    cache = getattr(owner, '_key_segments', None)
    if cache is None:
        cache = owner._key_segments = {}
    hit = cache.get(name)
    if hit is not None and hit[0] is keys and hit[1].nseg == nseg:
        return hit[1]
    k = keys.long()
    order = torch.argsort(k, stable=True).to(torch.int32).contiguous()
    ptr = torch.zeros(nseg + 1, dtype=torch.int32, device=keys.device)
    ptr[1:] = torch.bincount(k, minlength=nseg).cumsum(0).to(torch.int32)
    seg = _Segments(ptr, order, nseg, k.numel(), keys.device)
    cache[name] = (keys, seg)
    return seg


class GatherFew(torch.autograd.Function):
    """table[idx] for a table of FEW rows gathered MANY times (model_search_lp.py:171 `rel_embedding[triplets[:, 1]]`:
    330,000 scored triplets over 11 relations at C3).  ATen's index_put backward serialises on the duplicates (tens of
    milliseconds); here the backward is the reduction GEMM onehot(idx)^T . dy on the tensor cores (mrg_gemm_red: the
    one-hot entries are exact in tf32, so this is an fp32-class deterministic segmented sum)."""

    @staticmethod
    def forward(ctx, table, idx):
        table = _f32c(table)
        ctx.save_for_backward(idx)
        ctx.n_rows = table.shape[0]
        return table.index_select(0, idx)

    @staticmethod
    def backward(ctx, dy):
        idx, = ctx.saved_tensors
        dy = _f32c(dy)
        onehot = torch.zeros(idx.numel(), ctx.n_rows, dtype=torch.float32, device=dy.device)
        onehot.scatter_(1, idx.view(-1, 1).long(), 1.0)
        return gemm_red(onehot, dy), None


def gather_few(table, idx):
    """table[idx]; tables of at most 64 rows on a CUDA device take the GatherFew path."""
    if table.is_cuda and table.dim() == 2 and table.shape[0] <= 64 and idx.dim() == 1:
        return GatherFew.apply(table, idx)
    return table[idx]


# ------------------------------------------------------------------------------------------
# BatchNorm1d (+ReLU) over rows
# ------------------------------------------------------------------------------------------
class BNAct(torch.autograd.Function):
    """s = relu?(BN(y)) in one pass after a deterministic column-statistics reduction.
    Replaces nn.BatchNorm1d + ReLU after every op (model_lp.py:31-33, cell_lp.py:21,31-32).
    `stats` may carry partial sums already produced by y's producer kernel."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, training, momentum, eps, relu, stats):
        y = _f32c(y)
        rows, D = y.shape
        dev = y.device
        a = torch.empty(D, dtype=torch.float32, device=dev)
        b = torch.empty_like(a)
        if training:
            nparts = stats_nparts(rows)
            if stats is None:
                stats = _stats_buf(nparts, D, dev)
                call("mrg_colstats", act(y), rows, D, ptr(stats), stream())
            else:
                nparts = stats.numel() // (2 * D)
            mean = torch.empty_like(a)
            invstd = torch.empty_like(a)
            ctx.part = _dist.current()
            stats, nparts, rows_g = _dist.sync_stats(ctx.part, stats, nparts, 2 * D, rows)
            call("mrg_bn_finalize", ptr(stats), nparts, rows_g, D, ptr(gamma), ptr(beta), float(eps), float(momentum),
                 ptr(running_mean), ptr(running_var), ptr(mean), ptr(invstd), ptr(a), ptr(b), stream())
        else:
            invstd = torch.rsqrt(running_var + eps)
            mean = running_mean
            a = (gamma * invstd).contiguous()
            b = (beta - a * mean).contiguous()
        s = torch.empty_like(y)
        call("mrg_affine_act", act(y, a, b, relu), rows, D, ptr(s), stream())
        ctx.training, ctx.relu = training, relu
        ctx.save_for_backward(y, gamma, mean, invstd, a, b)
        return s

    @staticmethod
    def backward(ctx, ds):
        y, gamma, mean, invstd, a, b = ctx.saved_tensors
        ds = _f32c(ds)
        rows, D = y.shape
        dev = y.device
        yact = act(y, a, b, ctx.relu)
        dy = torch.empty_like(y)
        if ctx.training:
            nparts = stats_nparts(rows)
            bst = _stats_buf(nparts, D, dev)
            call("mrg_bn_bwd_reduce", ptr(ds), yact, rows, D, ptr(bst), stream())
            dgamma = torch.empty(D, dtype=torch.float32, device=dev)
            dbeta = torch.empty_like(dgamma)
            coef = torch.empty(3 * D, dtype=torch.float32, device=dev)
            bst, nparts, rows_g = _dist.sync_stats(getattr(ctx, 'part', None), bst, nparts, 2 * D, rows)
            call("mrg_bn_bwd_finalize", ptr(bst), nparts, rows_g, D, ptr(gamma), ptr(mean), ptr(invstd), ptr(dgamma),
                 ptr(dbeta), ptr(coef), stream())
            call("mrg_bn_bwd_apply", ptr(ds), yact, ptr(coef), rows, D, ptr(dy), 0, stream())
            dgamma, dbeta = _dist.unshare_param_grads(getattr(ctx, 'part', None), dgamma, dbeta)
        else:
            coef = torch.cat([torch.zeros(2 * D, device=dev), a]).contiguous()
            call("mrg_bn_bwd_apply", ptr(ds), yact, ptr(coef), rows, D, ptr(dy), 0, stream())
            # eval-mode parameter grads (rarely needed): plain reductions
            dz = dy / a
            dgamma = (dz * (y - mean) * invstd).sum(0)
            dbeta = dz.sum(0)
        return dy, dgamma, dbeta, None, None, None, None, None, None, None


def bn_momentum(bn):
    """Running-statistics update factor.  momentum=None (PyTorch: cumulative average 1/num_batches_tracked) would
    need a host read of a device counter inside the captured step; the reference only builds default
    nn.BatchNorm1d(D) (momentum 0.1), so it is refused rather than silently replaced by 0.1."""
    if bn.momentum is None:
        raise NotImplementedError("BatchNorm1d(momentum=None) (cumulative average) is not supported by the fused "
                                  "BatchNorm path; the reference uses the default momentum=0.1")
    return float(bn.momentum)


def bn_act(y, bn, relu=True, stats=None):
    """Apply an nn.BatchNorm1d module's parameters/buffers through the fused kernels."""
    training = bn.training or not bn.track_running_stats
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    momentum = bn_momentum(bn)
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    return BNAct.apply(y, bn.weight, bn.bias, rm, rv, training, momentum, bn.eps, relu, stats)


# ------------------------------------------------------------------------------------------
# collapse of the gates' Linear pair: (W_s, b_s, a_s) per segment -> v1 [S, D1], v2 [S, K-D1], c [S]
# ------------------------------------------------------------------------------------------
class CollapseGates(torch.autograd.Function):
    """v[s] = a_s.weight @ W_s.weight, c[s] = a_s.weight . W_s.bias for up to three (W, a) pairs in one kernel each
    way (mrg_gate_collapse_fwd/bwd): no non-linearity sits between W and a in the reference
    (operations_lp.py:319-320), so the [rows,K]x[K,D] GEMM of the gate collapses to one dot product per row.
    Inputs: D1, then (W.weight [D,K], W.bias [D], a.weight [1,D]) per segment.  Gradients land on those tensors."""

    @staticmethod
    def forward(ctx, D1, *wba):
        nseg = len(wba) // 3
        Ws = [_f32c(t) for t in wba[0::3]]
        bs = [(_f32c(t) if t is not None else None) for t in wba[1::3]]
        As = [_f32c(t) for t in wba[2::3]]
        D, Kdim = Ws[0].shape
        dev = Ws[0].device
        gp = _lib.MrgGateParams()
        for s_ in range(3):
            gp.W[s_] = Ws[s_].data_ptr() if s_ < nseg else None
            gp.b[s_] = bs[s_].data_ptr() if (s_ < nseg and bs[s_] is not None) else None
            gp.a[s_] = As[s_].data_ptr() if s_ < nseg else None
        v1 = torch.empty(nseg, D1, dtype=torch.float32, device=dev)
        v2 = torch.empty(nseg, Kdim - D1, dtype=torch.float32, device=dev) if Kdim > D1 else None
        c = torch.empty(nseg, dtype=torch.float32, device=dev)
        call("mrg_gate_collapse_fwd", gp, nseg, D, Kdim, D1, ptr(v1), ptr(v2), ptr(c), stream())
        ctx.dims = (nseg, D, Kdim, D1)
        ctx.has_b = [b is not None for b in bs]
        ctx.save_for_backward(*Ws, *[b for b in bs if b is not None], *As)
        if v2 is None:
            return v1, c
        return v1, v2, c

    @staticmethod
    def backward(ctx, *gouts):
        nseg, D, Kdim, D1 = ctx.dims
        saved = list(ctx.saved_tensors)
        Ws = saved[:nseg]
        nb = sum(ctx.has_b)
        b_saved = saved[nseg:nseg + nb]
        As = saved[nseg + nb:]
        bs, k = [], 0
        for hb in ctx.has_b:
            bs.append(b_saved[k] if hb else None)
            k += 1 if hb else 0
        if Kdim > D1:
            dv1, dv2, dc = gouts
        else:
            (dv1, dc), dv2 = gouts, None
        dev = Ws[0].device
        dv1 = _f32c(dv1) if dv1 is not None else torch.zeros(nseg, D1, dtype=torch.float32, device=dev)
        if Kdim > D1:
            dv2 = _f32c(dv2) if dv2 is not None else torch.zeros(nseg, Kdim - D1, dtype=torch.float32, device=dev)
        dc = _f32c(dc) if dc is not None else torch.zeros(nseg, dtype=torch.float32, device=dev)
        gp, gg = _lib.MrgGateParams(), _lib.MrgGateGrads()
        dWs = [torch.empty_like(W) for W in Ws]
        dbs = [torch.empty_like(b) if b is not None else None for b in bs]
        das = [torch.empty(1, D, dtype=torch.float32, device=dev) for _ in range(nseg)]
        for s_ in range(3):
            live = s_ < nseg
            gp.W[s_] = Ws[s_].data_ptr() if live else None
            gp.b[s_] = bs[s_].data_ptr() if (live and bs[s_] is not None) else None
            gp.a[s_] = As[s_].data_ptr() if live else None
            gg.dW[s_] = dWs[s_].data_ptr() if live else None
            gg.db[s_] = dbs[s_].data_ptr() if (live and dbs[s_] is not None) else None
            gg.da[s_] = das[s_].data_ptr() if live else None
        call("mrg_gate_collapse_bwd", gp, nseg, D, Kdim, D1, ptr(dv1), ptr(dv2), ptr(dc), gg, stream())
        out = [None]
        for s_ in range(nseg):
            out += [dWs[s_], dbs[s_], das[s_]]
        return tuple(out)


def collapse_gates(D1, pairs):
    """pairs: [(W Linear, a Linear)] per row segment -> (v1 [S,D1], v2 [S,K-D1] or None, c [S])."""
    flat = []
    for W, a in pairs:
        flat += [W.weight, W.bias, a.weight]
    res = CollapseGates.apply(D1, *flat)
    return res if len(res) == 3 else (res[0], None, res[1])


# ------------------------------------------------------------------------------------------
# K3/K6: collapsed sparse gate over one or more row segments
# ------------------------------------------------------------------------------------------
class SparseGate(torch.autograd.Function):
    """y = scale_i * sigmoid(x.v1 + xin.v2 + c) * x per row segment (v,c collapsed from the
    reference's W,a Linear pair) -- f_sparse_op_comp / f_sparse_op / f_sparse_op_last."""

    @staticmethod
    def forward(ctx, x, xin, v1, v2, c, bounds, row_scale, n_scaled, base_scales):
        x, xin = _f32c(x), _f32c(xin)
        v1, c = _f32c(v1), _f32c(c)
        v2 = _f32c(v2) if xin is not None else None
        rows, D = x.shape
        dev = x.device
        y = torch.empty_like(x)
        gate = torch.empty(rows, dtype=torch.float32, device=dev)
        nparts = [stats_nparts(hi - lo) for lo, hi in bounds]
        stats = _stats_buf(sum(nparts), D, dev)
        off = 0
        for s, (lo, hi) in enumerate(bounds):
            n = hi - lo
            rs = row_scale[lo:] if (row_scale is not None and lo < n_scaled) else None
            xa = act(x[lo:hi])
            ia = act(xin[lo:hi]) if xin is not None else act(None)
            call("mrg_sparse_gate_fwd", xa, ia, n, D, ptr(v1[s]), ptr(v2[s]) if v2 is not None else None,
                 ptr(c[s:s + 1]), ptr(rs), float(base_scales[s]), ptr(y[lo:hi]), ptr(gate[lo:hi]),
                 ptr(stats[off * 2 * D:]), stream())
            off += nparts[s]
        ctx.bounds, ctx.n_scaled, ctx.base_scales = bounds, n_scaled, base_scales
        ctx.has_in = xin is not None
        ctx.save_for_backward(x, xin, v1, v2, gate, row_scale)
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _):
        x, xin, v1, v2, gate, row_scale = ctx.saved_tensors
        dy = _f32c(dy)
        rows, D = x.shape
        dev = x.device
        dx = torch.empty_like(x)
        same = ctx.has_in and xin.data_ptr() == x.data_ptr()
        dxin = dx if same else (torch.empty_like(x) if ctx.has_in else None)
        dv1 = torch.empty_like(v1)
        dv2 = torch.empty_like(v2) if ctx.has_in else None
        dc = torch.empty(len(ctx.bounds), dtype=torch.float32, device=dev)
        lib = _lib.load()
        for s, (lo, hi) in enumerate(ctx.bounds):
            n = hi - lo
            rs = row_scale[lo:] if (row_scale is not None and lo < ctx.n_scaled) else None
            dparam = torch.empty(int(lib.mrg_gate_dparam_count(n, D)), dtype=torch.float64, device=dev)
            xa = act(x[lo:hi])
            ia = act(xin[lo:hi]) if ctx.has_in else act(None)
            call("mrg_sparse_gate_bwd", ptr(dy[lo:hi]), xa, ia, ptr(gate[lo:hi]), n, D, ptr(v1[s]),
                 ptr(v2[s]) if ctx.has_in else None, ptr(rs), float(ctx.base_scales[s]), ptr(dx[lo:hi]),
                 ptr(dxin[lo:hi]) if ctx.has_in else None, 0, ptr(dparam), stream())
            call("mrg_sparse_gate_bwd_finalize", ptr(dparam), n, D, ptr(dv1[s]), ptr(dv2[s]) if ctx.has_in else None,
                 ptr(dc[s:s + 1]), stream())
        if same:
            return dx, None, dv1, dv2, dc, None, None, None, None
        return dx, dxin, dv1, dv2, dc, None, None, None, None


class DenseGate(torch.autograd.Function):
    """y = scale_i * sigmoid?(z) * x : epilogue of f_dense_op_comp / f_comp_op / f_dense_op(_last)
    after the edge-tile GEMM z = W_s [x, xin] (+b)."""

    @staticmethod
    def forward(ctx, z, x, use_sigmoid, row_scale, n_scaled, base_scales, bounds):
        z, x = _f32c(z), _f32c(x)
        rows, D = z.shape
        y = torch.empty_like(z)
        nparts = [stats_nparts(hi - lo) for lo, hi in bounds]
        stats = _stats_buf(sum(nparts), D, z.device)
        off = 0
        for s, (lo, hi) in enumerate(bounds):
            rs = row_scale[lo:] if (row_scale is not None and lo < n_scaled) else None
            call("mrg_dense_gate_fwd", ptr(z[lo:hi]), act(x[lo:hi]) if use_sigmoid else act(None), hi - lo, D,
                 1 if use_sigmoid else 0, ptr(rs), float(base_scales[s]), ptr(y[lo:hi]), ptr(stats[off * 2 * D:]),
                 stream())
            off += nparts[s]
        ctx.meta = (use_sigmoid, n_scaled, base_scales, bounds)
        ctx.save_for_backward(z, x, row_scale)
        ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _):
        use_sigmoid, n_scaled, base_scales, bounds = ctx.meta
        z, x, row_scale = ctx.saved_tensors
        dy = _f32c(dy)
        D = z.shape[1]
        dz = torch.empty_like(z)
        dx = torch.empty_like(z) if use_sigmoid else None
        for s, (lo, hi) in enumerate(bounds):
            rs = row_scale[lo:] if (row_scale is not None and lo < n_scaled) else None
            call("mrg_dense_gate_bwd", ptr(dy[lo:hi]), ptr(z[lo:hi]), act(x[lo:hi]) if use_sigmoid else act(None),
                 hi - lo, D, 1 if use_sigmoid else 0, ptr(rs), float(base_scales[s]), ptr(dz[lo:hi]),
                 ptr(dx[lo:hi]) if use_sigmoid else None, 0, stream())
        return dz, dx, None, None, None, None, None


# ------------------------------------------------------------------------------------------
# K5: destination aggregation (DGL update_all replacement)
# ------------------------------------------------------------------------------------------
def decode_arg(arg):
    """Encoded argmax -> original edge ids (-1 = isolated node); see segreduce.cu."""
    return torch.where(arg < -1, -2 - arg, arg)


class SegReduce(torch.autograd.Function):
    """out[n] = REDUCE_{e: dst[e]=n} relu?(m[e]) (+ residual[n]) with REDUCE in sum|mean|max.
    max: isolated node -> 0, arg = lowest edge id attaining the max, gradient to that edge."""

    @staticmethod
    def forward(ctx, m, residual, g, kind, relu):
        m, residual = _f32c(m), _f32c(residual)
        E, D = m.shape
        N = g.N
        out = torch.empty(N, D, dtype=torch.float32, device=m.device)
        arg = torch.empty(N, D, dtype=torch.int32, device=m.device) if kind == 2 else None
        seg_reduce_raw(g.csr, kind, act(m, relu=relu), D, out, arg=arg,
                       residual=act(residual) if residual is not None else None)
        ctx.g, ctx.kind, ctx.relu, ctx.has_res = g, kind, relu, residual is not None
        need_m = relu and kind != 2
        ctx.save_for_backward(m if need_m else None, arg)
        g.last_arg = arg
        return out

    @staticmethod
    def backward(ctx, gout):
        g, kind = ctx.g, ctx.kind
        m, arg = ctx.saved_tensors
        gout = _f32c(gout)
        D = gout.shape[1]
        dm = torch.empty(g.E, D, dtype=torch.float32, device=gout.device)
        call("mrg_seg_reduce_bwd", kind, ptr(gout), ptr(arg), None, act(m, relu=ctx.relu) if m is not None else act(None),
             ptr(g.dst), ptr(g.csr.ptr), g.E, 0, D, ptr(dm), 0, stream())
        return dm, (gout if ctx.has_res else None), None, None, None


def amax_tc_supported(D):
    return bool(_lib.load().mrg_amax_tc_supported(int(D)))


AMAX_PRECISION = "fp32"     # "bf16": the reduced-precision variant of the fused a_max forward (mrg_amax_tc_fwd_bf16)


def amax_tc_call():
    if AMAX_PRECISION not in ("fp32", "bf16"):
        raise ValueError(f"AMAX_PRECISION must be 'fp32' or 'bf16', got {AMAX_PRECISION!r}")
    return "mrg_amax_tc_fwd_bf16" if AMAX_PRECISION == "bf16" else "mrg_amax_tc_fwd"


_tc_ws = {}


def _tc_workspace(N, D, device):
    key = (N, D, str(device))
    if key not in _tc_ws:
        nbytes = int(_lib.load().mrg_amax_tc_workspace_bytes(N, D))
        _tc_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return _tc_ws[key]


class AMaxTC(torch.autograd.Function):
    """a_max_op as ONE tcgen05 kernel: relu(W x_e + b) for the E edge rows, destination max with
    argmax, + residual rows; the [E,D] messages never reach HBM (operations_lp.py:230-235).
    x is [M,D] (LP: E edge rows then N self-loop rows, n_res=N) or [E,D] (NC blocks, n_res=0)."""

    @staticmethod
    def forward(ctx, x, weight, bias, g, has_residual):
        x, weight, bias = _f32c(x), _f32c(weight), _f32c(bias)
        D = x.shape[1]
        E, N = g.E, g.N
        out = torch.empty(N, D, dtype=torch.float32, device=x.device)
        arg = torch.empty(N, D, dtype=torch.int32, device=x.device)
        ws = _tc_workspace(N, D, x.device)
        res = act(x[E:]) if has_residual else act(None)
        call(amax_tc_call(), act(x), ptr(weight), ptr(bias), ptr(g.csr.idx), ptr(g.dst), E, N, D, res, ptr(out),
             ptr(arg), ptr(ws), ws.numel(), stream())
        ctx.g, ctx.has_residual = g, has_residual
        ctx.save_for_backward(x, weight, arg)
        g.last_arg = arg
        if getattr(g, 'arg_trace', None) is not None:
            g.arg_trace.append(arg)
        return out

    @staticmethod
    def backward(ctx, gout):
        g = ctx.g
        x, weight, arg = ctx.saved_tensors
        gout = _f32c(gout)
        dx, dw, db = amax_backward(g, gout, arg, act(x), weight, x.shape[0], ctx.has_residual, ctx.needs_input_grad[0])
        return dx, dw, db, None, None


_bwd_ws = {}


def amax_backward(g, gout, arg, x_act, weight, rows, has_residual, need_dx=True, dx=None):
    """Sparse backward of a_max (mrg_amax_bwd): dX over the E edge rows (+ residual rows = gout), dW, db."""
    D = gout.shape[1]
    dev = gout.device
    E = g.E
    key = (g.N, E, D, str(dev))
    if key not in _bwd_ws:
        _bwd_ws[key] = torch.empty(int(_lib.load().mrg_amax_bwd_workspace_bytes(g.N, E, D)), dtype=torch.uint8,
                                   device=dev)
    ws = _bwd_ws[key]
    if getattr(g, 'csr_dst', None) is None:      # destination of every dst-CSR position (graph-static), padded by
        pad = torch.zeros(64, dtype=torch.int32, device=dev)          # one 64-row window for the bulk copies
        g.csr_dst = torch.cat([g.dst[g.csr.idx[:E].long()], pad]).contiguous() if E > 0 else pad
    if need_dx and dx is None:
        dx = torch.empty(rows, D, dtype=torch.float32, device=dev)
    dw = torch.empty(D, D, dtype=torch.float32, device=dev)
    db = torch.empty(D, dtype=torch.float32, device=dev)
    call("mrg_amax_bwd", ptr(gout), ptr(arg), x_act, ptr(weight), ptr(g.csr.ptr), ptr(g.csr.idx), ptr(g.csr_dst),
         ptr(g.csr.chunk_first), ptr(g.csr.chunk_seg), g.N, E, g.csr.max_chunks, D, ptr(dx) if need_dx else None,
         ptr(dw), ptr(db), ptr(ws), ws.numel(), stream(), nbytes=2 * E * 4 * D + 2 * g.N * 4 * D)
    if need_dx and has_residual:
        dx[E:].copy_(gout)
    return (dx if need_dx else None), dw, db


class AggSumLP(torch.autograd.Function):
    """a_sum_op (LP, operations_lp.py:259-264) on the full [M,D] input: sum of the E edge rows
    by destination + the N self-loop rows; one backward kernel writes all M gradient rows."""

    @staticmethod
    def forward(ctx, x, g):
        x = _f32c(x)
        D = x.shape[1]
        out = torch.empty(g.N, D, dtype=torch.float32, device=x.device)
        seg_reduce_raw(g.csr, 0, act(x), D, out, residual=act(x[g.E:]))
        ctx.g = g
        return out

    @staticmethod
    def backward(ctx, gout):
        g = ctx.g
        gout = _f32c(gout)
        D = gout.shape[1]
        dx = torch.empty(g.M, D, dtype=torch.float32, device=gout.device)
        call("mrg_seg_reduce_bwd", 0, ptr(gout), None, None, act(None), ptr(g.dst), ptr(g.csr.ptr), g.E, g.N, D,
             ptr(dx), 0, stream())
        return dx, None


# ------------------------------------------------------------------------------------------
# K8: sigmoid + BCE over the 1-N logits
# ------------------------------------------------------------------------------------------
class SigmoidBCE(torch.autograd.Function):
    """mean BCE(sigmoid(logit), label) with torch's log clamp at -100 -- replaces
    torch.sigmoid (operations_lp.py:126) + nn.BCELoss (mr_lp_train.py:116,235)."""

    @staticmethod
    def forward(ctx, logit, label):
        logit, label = _f32c(logit), _f32c(label)
        n = logit.numel()
        nparts = int(_lib.load().mrg_bce_nparts(n))
        partial = torch.empty(nparts, dtype=torch.float64, device=logit.device)
        loss = torch.empty(1, dtype=torch.float32, device=logit.device)
        call("mrg_sigmoid_bce_fwd", ptr(logit), ptr(label), n, None, ptr(partial), ptr(loss), stream())
        ctx.save_for_backward(logit, label)
        return loss[0]

    @staticmethod
    def backward(ctx, gl):
        logit, label = ctx.saved_tensors
        dl = torch.empty_like(logit)
        gs = gl.reshape(1).float().contiguous()
        call("mrg_sigmoid_bce_bwd", ptr(logit), ptr(label), logit.numel(), ptr(gs), ptr(dl), stream())
        return dl, None


# ------------------------------------------------------------------------------------------
# node-level Linear on the tcgen05 main loop (3xTF32)
# ------------------------------------------------------------------------------------------
USE_TC_LINEAR = True
_lin_ws = {}


def _linear_tc(x, W, bias, out_cols=None):
    """out = x @ W.T (+ bias) through mrg_linear_tc_fwd; x [rows, K], W [F, K] contiguous fp32."""
    lib = _lib.load()
    rows, Kd = x.shape
    F_ = W.shape[0]
    key = (Kd, str(x.device))
    if key not in _lin_ws:
        _lin_ws[key] = torch.empty(int(lib.mrg_linear_tc_workspace_bytes(Kd)), dtype=torch.uint8, device=x.device)
    ws = _lin_ws[key]
    out = torch.empty(rows, F_, dtype=torch.float32, device=x.device)
    call("mrg_linear_tc_fwd", ptr(x), ptr(W), ptr(bias), rows, Kd, F_, ptr(out), F_, ptr(ws), ws.numel(), stream(),
         nbytes=rows * (Kd + F_) * 4)
    return out


_red_ws = {}
USE_TC_GEMM_RED = os.environ.get("MRG_GEMM_RED", "1") != "0"      # False: torch.mm (library GEMM) for the reductions over the rows
LINEAR_FWD_ON_GEMM_RED = os.environ.get("MRG_LINEAR_FWD_RED", "1") != "0"
LINEAR_DX_ON_GEMM_RED = os.environ.get("MRG_LINEAR_DX_RED", "1") != "0"
# Small dense table products (K.matmul: rel_wt @ embedding_e, rel_embed @ w_rel) on mrg_gemm_red: OFF by default.  They
# cost 10-16 us either way, but every edge of the graph reads one of the few rows of their result, so their rounding is
# coherent over all edges: 3xTF32's ~1e-6 (against ~1e-7 for an fp32 FMA GEMM) showed up 70x above the reference's own
# fp32 error in a cancellation-heavy BatchNorm bias gradient of the C3 supernet (test_c3_supernet_step_vs_real_reference).
USE_TC_MATMUL = os.environ.get("MRG_MATMUL_TC", "0") != "0"
# NC path: Linear on the relation TABLE, then gather (instead of gather, then Linear on every edge row), with the gather's
# backward as a segmented sum over the block's edge-type segments (model.py)
NC_REL_REORDER = os.environ.get("MRG_NC_REL_REORDER", "1") != "0"


def gemm_red(A, B, a_kmajor=False, colsum=False, bias=None):
    """A^T B with the reduction over the rows on the tensor cores (mrg_gemm_red, 3xTF32): A [rows, F1], B [rows, F2]
    -> [F1, F2]; a_kmajor: A is passed as At [F1, rows].  Weight gradients and the DistMult backward GEMMs.
    colsum: also return A.sum(0) (the bias gradient), computed by the same launch."""
    A, B = _f32c(A), _f32c(B)
    rows, F2 = B.shape
    F1 = A.shape[0] if a_kmajor else A.shape[1]
    if not (USE_TC_GEMM_RED and A.is_cuda and rows >= 64):
        C = torch.mm(A if a_kmajor else A.t(), B)
        if bias is not None:
            C = C + bias
        return (C, A.sum(1 if a_kmajor else 0)) if colsum else C
    lib = _lib.load()
    nbytes = int(lib.mrg_gemm_red_workspace_bytes(rows, F1, F2))
    key = str(A.device)
    if key not in _red_ws or _red_ws[key].numel() < nbytes:
        _red_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=A.device)
    ws = _red_ws[key]
    C = torch.empty(F1, F2, dtype=torch.float32, device=A.device)
    cs = torch.empty(F1, dtype=torch.float32, device=A.device) if colsum else None
    call("mrg_gemm_red", ptr(A), A.shape[1], 1 if a_kmajor else 0, ptr(B), F2, rows, F1, F2, ptr(C), F2, ptr(cs), ptr(bias),
         ptr(ws), ws.numel(), stream(), nbytes=rows * (F1 + F2) * 4)
    if nbytes > 16:
        _lib.launch_count += 1      # split reduction: main kernel + fold kernel (one slice: the main kernel alone)
    return (C, cs) if colsum else C


class MatmulTC(torch.autograd.Function):
    """X @ Y for small dense tables on mrg_gemm_red (X [m, k] is its own K-major operand, Y [k, n] the row-major one):
    `rel_wt @ embedding_e.weight` and `rel_embed @ w_rel` (model_lp.py:125,133), whose inner dimension (2R+1 = 475) rules
    out the K % 8 == 0 Linear kernel.  Backward: dX = dC Y^T and dY = X^T dC on the same entry point."""

    @staticmethod
    def forward(ctx, X, Y):
        X, Y = _f32c(X), _f32c(Y)
        ctx.save_for_backward(X, Y)
        return gemm_red(X, Y, a_kmajor=True)

    @staticmethod
    def backward(ctx, dC):
        X, Y = ctx.saved_tensors
        dC = _f32c(dC)
        dX = gemm_red(dC, Y.t().contiguous(), a_kmajor=True) if ctx.needs_input_grad[0] else None
        dY = gemm_red(X, dC) if ctx.needs_input_grad[1] else None
        return dX, dY


def matmul(X, Y):
    """X @ Y (2-D, fp32) on the tensor cores when both are CUDA tensors, else torch.mm."""
    if USE_TC_GEMM_RED and USE_TC_MATMUL and X.is_cuda and X.dim() == 2 and Y.dim() == 2:
        return MatmulTC.apply(X, Y)
    return torch.mm(X, Y)


class LinearTC(torch.autograd.Function):
    """nn.Linear forward, input gradient and weight gradient on the tensor cores with fp32-class accuracy (3xTF32):
    the cell's `concat` Linear and `linear_e` (model_lp.py:70-71,124), the Linear candidates."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, weight = _f32c(x), _f32c(weight)
        bias = _f32c(bias) if bias is not None else None
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        if USE_TC_GEMM_RED and LINEAR_FWD_ON_GEMM_RED and x.shape[1] >= 128:
            # x [rows, K] is its own K-major operand.  (Short reductions, e.g. the D = 64 Linear on 12 M edge rows of the NC
            # path, keep the persistent kernel: a 256-row CTA tile would do two K chunks of work per launch overhead.)
            return gemm_red(x, weight.t().contiguous(), a_kmajor=True, bias=bias)
        return _linear_tc(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy = _f32c(gy)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if USE_TC_GEMM_RED and LINEAR_DX_ON_GEMM_RED and gy.shape[1] >= 128:
                dx = gemm_red(gy, weight, a_kmajor=True)            # dY [rows, F] is its own K-major operand, W [F, K] row-major
            else:
                dx = _linear_tc(gy, weight.t().contiguous(), None)      # dY [rows, F] x W [F, K]
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1]:
            dw = gemm_red(gy, x, colsum=want_db)
            if want_db:
                dw, db = dw
        elif want_db:
            db = gy.sum(0)
        return dx, dw, db


def linear_fn(x, weight, bias):
    """x @ weight.T (+ bias): tcgen05 (3xTF32) for large tiles, the library GEMM for small ones."""
    if USE_TC_LINEAR and x.is_cuda and x.dim() == 2 and x.shape[1] % 8 == 0 and weight.shape[0] % 8 == 0 \
            and x.shape[0] >= 1024:
        return LinearTC.apply(x, weight, bias)
    return torch.nn.functional.linear(x, weight, bias)


def linear(module, x):
    """module(x) for an nn.Linear, on the tensor-core path when the shapes allow it."""
    if USE_TC_LINEAR and x.is_cuda and x.dim() == 2 and x.shape[1] % 8 == 0 and module.weight.shape[0] % 8 == 0 \
            and x.shape[0] >= 1024:
        return LinearTC.apply(x, module.weight, module.bias)
    return module(x)


_dm_ws = {}


class DistMultBCE(torch.autograd.Function):
    """loss = BCELoss(sigmoid((sub_emb * rel_emb) @ all_ent.T), label) in ONE tcgen05 kernel (mrg_distmult_bce_fwd):
    sf_DisMult_op.forward (operations_lp.py:115-127) + nn.BCELoss (mr_lp_train.py:116) for the training loss.
    Backward: dlogit from the stored logits (mrg_sigmoid_bce_bwd), then the two reduction GEMMs (mrg_gemm_red)."""

    @staticmethod
    def forward(ctx, all_ent, sub_emb, rel_emb, label):
        all_ent, sub_emb, rel_emb, label = _f32c(all_ent), _f32c(sub_emb), _f32c(rel_emb), _f32c(label)
        lib = _lib.load()
        B, D = sub_emb.shape
        N = all_ent.shape[0]
        dev = all_ent.device
        query = (sub_emb * rel_emb).contiguous()
        key = (D, str(dev))
        if key not in _dm_ws:
            _dm_ws[key] = torch.empty(int(lib.mrg_distmult_bce_workspace_bytes(D)), dtype=torch.uint8, device=dev)
        ws = _dm_ws[key]
        logit = torch.empty(B, N, dtype=torch.float32, device=dev)
        partial = torch.empty(int(lib.mrg_distmult_bce_nparts(B)), dtype=torch.float64, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        call("mrg_distmult_bce_fwd", ptr(query), ptr(all_ent), ptr(label), B, N, D, ptr(logit), ptr(partial), ptr(loss),
             ptr(ws), ws.numel(), stream(), nbytes=N * D * 4 + 2 * B * N * 4)
        ctx.save_for_backward(logit, label, all_ent, sub_emb, rel_emb, query)
        return loss[0]

    @staticmethod
    def backward(ctx, gl):
        logit, label, all_ent, sub_emb, rel_emb, query = ctx.saved_tensors
        dl = torch.empty_like(logit)
        gs = gl.reshape(1).float().contiguous()
        call("mrg_sigmoid_bce_bwd", ptr(logit), ptr(label), logit.numel(), ptr(gs), ptr(dl), stream())
        dq = gemm_red(dl, all_ent, a_kmajor=True)
        dent = gemm_red(dl, query) if ctx.needs_input_grad[0] else None
        return dent, dq * rel_emb, dq * sub_emb, None


_te_ws = {}


class TransELogits(torch.autograd.Function):
    """logit[b,n] = gamma - ||query[b] - all_ent[n]||_1 (sf_TransE_op, operations_lp.py:101-112) without the
    reference's [B, N, D] broadcast: mrg_transe_fwd / mrg_transe_bwd."""

    @staticmethod
    def forward(ctx, all_ent, query, gamma):
        all_ent, query = _f32c(all_ent), _f32c(query)
        B, D = query.shape
        N = all_ent.shape[0]
        logit = torch.empty(B, N, dtype=torch.float32, device=all_ent.device)
        call("mrg_transe_fwd", ptr(query), ptr(all_ent), B, N, D, float(gamma), ptr(logit), stream(),
             nbytes=N * D * 4 + B * N * 4)
        ctx.save_for_backward(all_ent, query)
        return logit

    @staticmethod
    def backward(ctx, dl):
        all_ent, query = ctx.saved_tensors
        dl = _f32c(dl)
        B, D = query.shape
        N = all_ent.shape[0]
        dev = dl.device
        key = (B, N, D, str(dev))
        if key not in _te_ws:
            _te_ws[key] = torch.empty(int(_lib.load().mrg_transe_bwd_workspace_bytes(B, N, D)), dtype=torch.uint8,
                                      device=dev)
        ws = _te_ws[key]
        dq = torch.empty_like(query) if ctx.needs_input_grad[1] else None
        de = torch.empty_like(all_ent) if ctx.needs_input_grad[0] else None
        call("mrg_transe_bwd", ptr(dl), ptr(query), ptr(all_ent), B, N, D, ptr(dq), ptr(de), ptr(ws), ws.numel(),
             stream(), nbytes=2 * B * N * 4)
        return de, dq, None


def distmult_bce_supported(D):
    return bool(_lib.load().mrg_distmult_bce_supported(int(D)))


# ------------------------------------------------------------------------------------------
# plain ReLU through the same row kernels (NC OpModule without op_norm: model.py:22-28)
# ------------------------------------------------------------------------------------------
class ReluAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y):
        y = _f32c(y)
        rows, D = y.shape
        s = torch.empty_like(y)
        call("mrg_affine_act", act(y, relu=True), rows, D, ptr(s), stream())
        ctx.save_for_backward(y)
        return s

    @staticmethod
    def backward(ctx, ds):
        (y,) = ctx.saved_tensors
        ds = _f32c(ds)
        rows, D = y.shape
        coef = torch.cat([torch.zeros(2 * D, device=y.device), torch.ones(D, device=y.device)]).contiguous()
        dy = torch.empty_like(y)
        call("mrg_bn_bwd_apply", ptr(ds), act(y, relu=True), ptr(coef), rows, D, ptr(dy), 0, stream())
        return dy


# ------------------------------------------------------------------------------------------
# K9: DARTS MixedOp = sum_k w_k * ReLU(BN_k(y_k)) in one pass
# ------------------------------------------------------------------------------------------
class MixedSum(torch.autograd.Function):
    """cell_lp.py:25-33 / cell.py:23-31.  Inputs: w [K] (a softmax(alpha) row), then per candidate
    (y_k, gamma_k, beta_k).  BN statistics, normalisation, ReLU and the weighted sum are fused; the
    backward folds w_k into each candidate's BatchNorm backward (no [rows,D] temporaries per candidate
    besides the returned dy_k).  y_k = None stands for an identically-zero candidate (f_zero_op): its BatchNorm sees
    zero statistics, its output relu(beta) is added from the affine alone and nothing of it is ever materialised."""

    @staticmethod
    def forward(ctx, w, training, eps, momentum, bns, stats_list, like, *tensors):
        K_ = len(tensors) // 3
        ys = [_f32c(t) if t is not None else None for t in tensors[0::3]]
        gammas, betas = tensors[1::3], tensors[2::3]
        rows, D = like.shape
        dev = like.device
        w = _f32c(w)
        saved = []
        lst = _lib.MrgActList()
        lst.n = K_
        for k in range(K_):
            a = torch.empty(D, dtype=torch.float32, device=dev)
            b = torch.empty_like(a)
            rm, rv = bns[k]
            if training:
                st = stats_list[k]
                nparts = stats_nparts(rows)
                if ys[k] is None:
                    st, nparts = torch.zeros(1, 2, D, dtype=torch.float64, device=dev), 1
                elif st is None:
                    st = _stats_buf(nparts, D, dev)
                    call("mrg_colstats", act(ys[k]), rows, D, ptr(st), stream())
                else:
                    nparts = st.numel() // (2 * D)
                mean, invstd = torch.empty_like(a), torch.empty_like(a)
                call("mrg_bn_finalize", ptr(st), nparts, rows, D, ptr(gammas[k]), ptr(betas[k]), float(eps),
                     float(momentum), ptr(rm), ptr(rv), ptr(mean), ptr(invstd), ptr(a), ptr(b), stream())
            else:
                invstd = torch.rsqrt(rv + eps)
                mean = rm
                a = (gammas[k] * invstd).contiguous()
                b = (betas[k] - a * mean).contiguous()
            lst.acts[k] = act(ys[k], a, b, True) if ys[k] is not None else _lib.MrgAct(None, a.data_ptr(), b.data_ptr(), 1)
            saved += [ys[k], gammas[k], mean, invstd, a, b]
        out = torch.empty(rows, D, dtype=torch.float32, device=dev)
        call("mrg_mixed_sum_fwd", lst, ptr(w), rows, D, ptr(out), stream())
        ctx.K_, ctx.training = K_, training
        ctx.save_for_backward(w, *saved)
        return out

    @staticmethod
    def backward(ctx, dout):
        w, *saved = ctx.saved_tensors
        K_ = ctx.K_
        dout = _f32c(dout)
        rows, D = dout.shape
        dev = dout.device
        dw = torch.empty(K_, dtype=torch.float32, device=dev)
        grads = []
        nparts = stats_nparts(rows)
        for k in range(K_):
            y, gamma, mean, invstd, a, b = saved[6 * k:6 * k + 6]
            yact = act(y, a, b, True) if y is not None else _lib.MrgAct(None, a.data_ptr(), b.data_ptr(), 1)
            bst = _stats_buf(nparts, D, dev)
            call("mrg_bn_bwd_reduce", ptr(dout), yact, rows, D, ptr(bst), stream())
            dgamma = torch.empty(D, dtype=torch.float32, device=dev)
            dbeta = torch.empty_like(dgamma)
            coef = torch.empty(3 * D, dtype=torch.float32, device=dev)
            call("mrg_bn_bwd_finalize", ptr(bst), nparts, rows, D, ptr(gamma), ptr(mean), ptr(invstd), ptr(dgamma),
                 ptr(dbeta), ptr(coef), stream())
            # dw[k] = sum(dout * s_k) = sum_c a_c * S2_c + b_c * S1_c (S1 = dbeta, S2 = dgamma/invstd + mean*S1), then
            # coef / dgamma / dbeta scaled by w[k] (eval: dy = w_k * a * dz) -- one launch
            call("mrg_mixed_bwd_scale", ptr(coef), ptr(dgamma), ptr(dbeta), ptr(a), ptr(b), ptr(mean), ptr(invstd),
                 ptr(w), k, ptr(dw), D, 1 if ctx.training else 0, stream())
            dy = None
            if y is not None:
                dy = torch.empty_like(y)
                call("mrg_bn_bwd_apply", ptr(dout), yact, ptr(coef), rows, D, ptr(dy), 0, stream())
            grads += [dy, dgamma, dbeta]
        return (dw, None, None, None, None, None, None, *grads)


class MixedPre(torch.autograd.Function):
    """MixedOp over composition candidates (PRE_OPS) with one shared read of (a, b): cell_lp.py:25-33 applied to
    pre_mult / pre_sub / pre_add (operations_lp.py:71-98).  No candidate output is materialised (mixed_pre.cu):
    statistics pass -> K BatchNorm finalizes -> mixed-sum pass; the backward mirrors it."""

    @staticmethod
    def forward(ctx, w, training, eps, momentum, bns, comps, a, b, *gamma_beta):
        a, b, w = _f32c(a), _f32c(b), _f32c(w)
        rows, D = a.shape
        dev = a.device
        K_ = len(comps)
        gammas, betas = gamma_beta[0::2], gamma_beta[1::2]
        comps_t = torch.tensor(list(comps), dtype=torch.int32)      # host array read at enqueue time
        ctx.comps_t = comps_t
        cp = ctypes.c_void_p(comps_t.data_ptr())
        scale = torch.empty(K_, D, dtype=torch.float32, device=dev)
        shift = torch.empty_like(scale)
        mean = torch.empty_like(scale)
        invstd = torch.empty_like(scale)
        if training:
            nparts = stats_nparts(rows)
            st = torch.empty(K_, nparts, 2, D, dtype=torch.float64, device=dev)
            call("mrg_mixed_pre_stats", ptr(a), ptr(b), rows, D, cp, K_, ptr(st), stream(), nbytes=2 * rows * D * 4)
            for k in range(K_):
                rm, rv = bns[k]
                call("mrg_bn_finalize", ptr(st[k]), nparts, rows, D, ptr(gammas[k]), ptr(betas[k]), float(eps),
                     float(momentum), ptr(rm), ptr(rv), ptr(mean[k]), ptr(invstd[k]), ptr(scale[k]), ptr(shift[k]), stream())
        else:
            for k in range(K_):
                rm, rv = bns[k]
                invstd[k] = torch.rsqrt(rv + eps)
                mean[k] = rm
                scale[k] = gammas[k] * invstd[k]
                shift[k] = betas[k] - scale[k] * rm
        out = torch.empty_like(a)
        call("mrg_mixed_pre_fwd", ptr(a), ptr(b), rows, D, cp, K_, ptr(scale), ptr(shift), ptr(w), ptr(out), stream(),
             nbytes=3 * rows * D * 4)
        ctx.K_, ctx.training = K_, training
        ctx.save_for_backward(a, b, w, scale, shift, mean, invstd, *gammas)
        return out

    @staticmethod
    def backward(ctx, dout):
        a, b, w, scale, shift, mean, invstd, *gammas = ctx.saved_tensors
        K_ = ctx.K_
        dout = _f32c(dout)
        rows, D = dout.shape
        dev = dout.device
        cp = ctypes.c_void_p(ctx.comps_t.data_ptr())
        nparts = stats_nparts(rows)
        bst = torch.empty(K_, nparts, 2, D, dtype=torch.float64, device=dev)
        call("mrg_mixed_pre_bwd_stats", ptr(dout), ptr(a), ptr(b), rows, D, cp, K_, ptr(scale), ptr(shift), ptr(bst),
             stream(), nbytes=3 * rows * D * 4)
        coef = torch.empty(K_, 3 * D, dtype=torch.float32, device=dev)
        dw = torch.empty(K_, dtype=torch.float32, device=dev)
        grads = []
        for k in range(K_):
            dgamma = torch.empty(D, dtype=torch.float32, device=dev)
            dbeta = torch.empty_like(dgamma)
            call("mrg_bn_bwd_finalize", ptr(bst[k]), nparts, rows, D, ptr(gammas[k]), ptr(mean[k]), ptr(invstd[k]),
                 ptr(dgamma), ptr(dbeta), ptr(coef[k]), stream())
            call("mrg_mixed_bwd_scale", ptr(coef[k]), ptr(dgamma), ptr(dbeta), ptr(scale[k]), ptr(shift[k]), ptr(mean[k]),
                 ptr(invstd[k]), ptr(w), k, ptr(dw), D, 1 if ctx.training else 0, stream())
            grads += [dgamma, dbeta]
        need_a, need_b = ctx.needs_input_grad[6], ctx.needs_input_grad[7]
        da = torch.empty_like(a) if need_a else None
        db = torch.empty_like(b) if need_b else None
        if need_a or need_b:
            call("mrg_mixed_pre_bwd", ptr(dout), ptr(a), ptr(b), rows, D, cp, K_, ptr(scale), ptr(shift), ptr(coef), ptr(da),
                 ptr(db), stream(), nbytes=5 * rows * D * 4)
        return (dw, None, None, None, None, None, da, db, *grads)


MIXED_PRE_FUSED = os.environ.get("MRG_MIXED_PRE", "1") != "0"      # False: per-candidate outputs + mixed_sum (the round-1 form)


def mixed_pre(weights, a, b, comps, bn_modules):
    """sum_k weights[k] * ReLU(bn_k(comp_k(a, b))) without materialising the candidates (MixedPre)."""
    training = bn_modules[0].training
    bns, flat = [], []
    for bn in bn_modules:
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        bns.append((bn.running_mean, bn.running_var))
        flat += [bn.weight, bn.bias]
    return MixedPre.apply(weights, training, bn_modules[0].eps, bn_momentum(bn_modules[0]), bns, tuple(comps), a, b, *flat)


def mixed_sum(weights, ys, bn_modules, like=None):
    """sum_k weights[k] * ReLU(bn_k(ys[k])) with nn.BatchNorm1d modules supplying parameters/buffers.  ys[k] may be None
    (an identically-zero candidate); `like` then supplies shape and device."""
    training = bn_modules[0].training
    bns, flat, stats = [], [], []
    if like is None:
        like = next(y for y in ys if y is not None)
    for y, bn in zip(ys, bn_modules):
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        bns.append((bn.running_mean, bn.running_var))
        stats.append(getattr(y, 'mrg_stats', None) if y is not None else None)
        flat += [y, bn.weight, bn.bias]
    mom = bn_momentum(bn_modules[0])
    return MixedSum.apply(weights, training, bn_modules[0].eps, mom, bns, stats, like.detach(), *flat)
