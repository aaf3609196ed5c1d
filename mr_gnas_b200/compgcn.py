"""CompGCN layer (reference: models/compgcn.py:12-113, an orphan mirror of DGL's example) on libmrgnas.

`sub` composition -- NODE-LEVEL form (north_star: "for sub, hW - rW is precomputed before the per-edge reduce").
The reference transforms every edge row: W_d (h_u - n_e r_t) + b_d for direction d in {out, in} (two masked
[E/2, D_in] x [D_in, D_out] GEMMs, an index_put and a DGL sum).  Linearity gives, exactly up to rounding,
    agg[v] = sum_{e: u->v, dir d} ( (h W_d^T)[u] - n_e (r W_d^T)[t_e] + b_d )
           = segsum_v( HW[2u + d] )  -  (A @ RW)[v]  +  (C @ [b_O; b_I])[v]
with HW = h [W_O; W_I]^T viewed [2N, D_out] (ONE node-level tcgen05 GEMM), RW = r [W_O; W_I]^T viewed
[2R', D_out], and two graph-static matrices: A[v, 2t + d] = sum of the norms of v's in-edges of type t and direction
d, C[v, d] = their count.  No [E, D] tensor is ever formed: one gather-sum over the dst-CSR reads HW rows by a
composite index (mrg_seg_reduce_fwd), its backward is the same kernel over the (source, direction) segments.
`mul` needs the product per edge and keeps the edge-level path (composition kernel + masked edge-tile GEMMs +
segmented sum); ``ccorr`` (circular correlation) uses torch.fft because the reference's torch.rfft
(utils/utils.py:301) no longer exists."""
import torch
import torch.nn as nn

from . import functional as K
from .graph import _Segments

NODE_LEVEL_SUB = True      # tests switch this off to compare the two forms
_MAX_DENSE_A = 1 << 26     # elements of the dense [N, 2R'] norm matrix above which the edge-level path is used


class _SegView:
    """The segment list of `base` (ptr, chunk tables, workspaces) read through another gather index."""

    def __init__(self, base, idx):
        self.ptr, self.idx, self.nseg, self.total = base.ptr, idx, base.nseg, base.total
        self.max_chunks, self.chunk_first, self.chunk_seg = base.max_chunks, base.chunk_first, base.chunk_seg
        self.workspace = base.workspace


class _GatherSegSum(torch.autograd.Function):
    """out[s] = sum_{p in segment s} table[fwd.idx[p]]; backward: dtable[j] = sum_{q in segment j of bwd} dout[bwd.idx[q]]
    (both mrg_seg_reduce_fwd: deterministic, no atomics)."""

    @staticmethod
    def forward(ctx, table, fwd, bwd, n_out):
        table = K._f32c(table)
        D = table.shape[1]
        out = torch.empty(n_out, D, dtype=torch.float32, device=table.device)
        K.seg_reduce_raw(fwd, 0, K.act(table), D, out)
        ctx.bwd, ctx.rows = bwd, table.shape[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = K._f32c(dout)
        D = dout.shape[1]
        dtable = torch.empty(ctx.rows, D, dtype=torch.float32, device=dout.device)
        K.seg_reduce_raw(ctx.bwd, 0, K.act(dout), D, dtable)
        return dtable, None, None, None


def _static_parts(g, etype, n_rel_rows):
    """Graph-static pieces of the node-level form, cached on the graph: composite gather indices, the
    (source, direction) segments for the backward, the dense norm matrix A and the count matrix C."""
    key = ("compgcn_static", int(n_rel_rows), g.edata['in_edges_mask'].data_ptr(), g.edata['out_edges_mask'].data_ptr())
    cache = g.__dict__.setdefault("_compgcn_cache", {})
    if key in cache:
        return cache[key]
    dev = g.src.device
    E, N = g.E, g.N
    src, dst = g.src.long(), g.dst.long()
    out_m, in_m = g.edata['out_edges_mask'].bool(), g.edata['in_edges_mask'].bool()
    d = torch.where(in_m, torch.ones_like(src), torch.zeros_like(src))      # later assignment wins, as index_put does
    live = out_m | in_m
    zero_row = 2 * N                                                          # edges in neither mask read a zero row
    idx1 = torch.where(live, 2 * src + d, torch.full_like(src, zero_row))
    eid = g.csr.idx[:E].long()
    fwd = _SegView(g.csr, idx1[eid].to(torch.int32).contiguous())
    order = torch.argsort(idx1, stable=True)
    ptr = torch.zeros(2 * N + 2, dtype=torch.int64, device=dev)
    ptr[1:] = torch.cumsum(torch.bincount(idx1, minlength=2 * N + 1), 0)
    bwd = _Segments(ptr.to(torch.int32), dst[order].to(torch.int32).contiguous(), 2 * N + 1, E, dev)
    norm = g.edata['norm'].reshape(-1).float()
    col = 2 * etype.long() + d
    A = torch.zeros(N, 2 * n_rel_rows, dtype=torch.float32, device=dev)
    A.index_put_((dst[live], col[live]), norm[live], accumulate=True)        # sort-based on CUDA: deterministic
    C = torch.zeros(N, 2, dtype=torch.float32, device=dev)
    C.index_put_((dst[live], d[live]), torch.ones_like(norm[live]), accumulate=True)
    cache[key] = (fwd, bwd, A, C)
    return cache[key]


def ccorr(a, b):
    n = a.shape[-1]
    return torch.fft.irfft(torch.conj(torch.fft.rfft(a, dim=-1)) * torch.fft.rfft(b, dim=-1), n=n, dim=-1)


class CompGraphConv(nn.Module):
    def __init__(self, in_dim, out_dim, comp_fn='sub', batchnorm=True, dropout=0.1):
        super().__init__()
        self.in_dim, self.out_dim, self.comp_fn = in_dim, out_dim, comp_fn
        self.actvation = torch.tanh
        self.batchnorm = batchnorm
        self.dropout = nn.Dropout(dropout)
        if self.batchnorm:
            self.bn = nn.BatchNorm1d(out_dim)
        self.W_O = nn.Linear(in_dim, out_dim)
        self.W_I = nn.Linear(in_dim, out_dim)
        self.W_S = nn.Linear(in_dim, out_dim)
        self.W_R = nn.Linear(in_dim, out_dim)
        self.loop_rel = nn.Parameter(torch.Tensor(1, in_dim))
        nn.init.xavier_normal_(self.loop_rel)

    def _comp(self, a, b):
        if self.comp_fn == 'sub':
            return K.ComposeRows.apply(a, b, 0)
        if self.comp_fn == 'mul':
            return K.ComposeRows.apply(a, b, 1)
        if self.comp_fn == 'ccorr':
            return ccorr(a, b)
        raise Exception('Only supports sub, mul, and ccorr')

    def forward(self, g, n_in_feats, r_feats):
        """g: MRGraph with edata 'etype' (or 'e_type'), 'norm', 'in_edges_mask', 'out_edges_mask'."""
        r_feats = torch.cat((r_feats, self.loop_rel), 0)
        etype = g.edata['etype'] if 'etype' in g.edata else g.edata['e_type']
        if (self.comp_fn == 'sub' and NODE_LEVEL_SUB and n_in_feats.is_cuda
                and g.N * 2 * r_feats.shape[0] <= _MAX_DENSE_A):
            comp_edge = self._agg_sub_node_level(g, n_in_feats, r_feats, etype)
            return self._tail(g, n_in_feats, r_feats, comp_edge)
        src, _ = g.edges()
        e_h = (r_feats[etype.long()] * g.edata['norm'].view(-1, 1)).contiguous()
        comp_h = self._comp(n_in_feats[src].contiguous(), e_h)
        in_idx = torch.nonzero(g.edata['in_edges_mask'], as_tuple=False).view(-1)
        out_idx = torch.nonzero(g.edata['out_edges_mask'], as_tuple=False).view(-1)
        new_comp_h = torch.zeros(comp_h.shape[0], self.out_dim, device=comp_h.device)
        new_comp_h = new_comp_h.index_put((out_idx,), self.W_O(comp_h[out_idx]))
        new_comp_h = new_comp_h.index_put((in_idx,), self.W_I(comp_h[in_idx]))
        comp_edge = K.SegReduce.apply(new_comp_h, None, g, 0, False)   # update_all(copy_e, sum)
        return self._tail(g, n_in_feats, r_feats, comp_edge)

    def _agg_sub_node_level(self, g, h, r_feats, etype):
        """sum over in-edges of W_d (h_u - n_e r_t) + b_d without any edge-level tensor (module docstring)."""
        fwd, bwd, A, C = _static_parts(g, etype, r_feats.shape[0])
        N, Dout = g.N, self.out_dim
        Wcat = torch.cat([self.W_O.weight, self.W_I.weight], 0)                 # [2 D_out, D_in]
        hw = K.linear_fn(h, Wcat, None)                                         # node-level GEMM (tcgen05 when large)
        table = torch.cat([hw.reshape(2 * N, Dout), hw.new_zeros(1, Dout)], 0)  # rows 2u + d, + the zero row
        term1 = _GatherSegSum.apply(table, fwd, bwd, N)
        rw = torch.mm(r_feats, Wcat.t()).reshape(2 * r_feats.shape[0], Dout)    # rows 2t + d
        term2 = K.linear_fn(A, rw.t().contiguous(), None)                       # A @ RW: node-level GEMM
        bias = torch.mm(C, torch.stack([self.W_O.bias, self.W_I.bias], 0))
        return term1 - term2 + bias

    def _tail(self, g, n_in_feats, r_feats, comp_edge):
        loop = r_feats[-1].expand_as(n_in_feats).contiguous()
        comp_h_s = self._comp(n_in_feats.contiguous(), loop)
        n_out = (self.W_S(comp_h_s) + self.dropout(comp_edge)) * (1 / 3)
        r_out = self.W_R(r_feats)
        if self.batchnorm:
            n_out = K.bn_act(n_out, self.bn, relu=False)
        if self.actvation is not None:
            n_out = self.actvation(n_out)
        return n_out, r_out[:-1]


class CompGCN(nn.Module):
    """reference: compgcn.py:116-185 (stack of CompGraphConv over basis-composed relation embeddings)."""

    def __init__(self, num_bases, num_rel, num_ent, in_dim=100, layer_size=(200,), comp_fn='sub', batchnorm=True,
                 dropout=0.1, layer_dropout=(0.3,)):
        super().__init__()
        self.num_bases, self.num_rel, self.num_ent = num_bases, num_rel, num_ent
        self.in_dim, self.layer_size, self.comp_fn = in_dim, list(layer_size), comp_fn
        self.num_layer = len(self.layer_size)
        self.layers = nn.ModuleList()
        self.layers.append(CompGraphConv(in_dim, self.layer_size[0], comp_fn, batchnorm, dropout))
        for i in range(self.num_layer - 1):
            self.layers.append(CompGraphConv(self.layer_size[i], self.layer_size[i + 1], comp_fn, batchnorm, dropout))
        if self.num_bases > 0:
            self.basis = nn.Parameter(torch.Tensor(self.num_bases, self.in_dim))
            self.weights = nn.Parameter(torch.Tensor(self.num_rel, self.num_bases))
            nn.init.xavier_normal_(self.basis)
            nn.init.xavier_normal_(self.weights)
        else:
            self.rel_embds = nn.Parameter(torch.Tensor(self.num_rel, self.in_dim))
            nn.init.xavier_normal_(self.rel_embds)
        self.n_embds = nn.Parameter(torch.Tensor(self.num_ent, self.in_dim))
        nn.init.xavier_normal_(self.n_embds)
        self.dropouts = nn.ModuleList([nn.Dropout(p) for p in layer_dropout])

    def forward(self, graph):
        n_feats = self.n_embds
        r_feats = torch.mm(self.weights, self.basis) if self.num_bases > 0 else self.rel_embds
        for layer, dropout in zip(self.layers, self.dropouts):
            n_feats, r_feats = layer(graph, n_feats, r_feats)
            n_feats = dropout(n_feats)
        return n_feats, r_feats
