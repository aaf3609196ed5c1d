"""Link-prediction fine-grained operators: same registry names, constructor dicts, parameter
names/shapes/registration order and ``forward(g, src_emb, src_emb_in)`` signatures as the
reference's models/operations_lp.py, executed by the libmrgnas sm_100a kernels.

Row layout of every edge-level tensor (reference contract): rows [0,E/2) original-direction
edges, [E/2,E) inverse edges, [E,E+N) self loops; row i <-> edge id i.

Outputs of ops that are followed by BatchNorm in the cells carry ``.mrg_stats`` (per-block
column sum / sum-of-squares partials written by the producing kernel) so the BN that follows
does not re-read the tensor.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as K

MIXED_OPS = {
    'pre_mult': lambda args: pre_mult_op(),
    'pre_sub': lambda args: pre_sub_op(),
    'pre_add': lambda args: pre_add_op(),
    'f_zero': lambda args: f_zero_op(),
    'f_identity': lambda args: f_identity_op(),
    'f_dense': lambda args: f_dense_op(args),
    'f_dense_comp': lambda args: f_dense_op_comp(args),
    'f_comp': lambda args: f_comp_op(args),
    'f_sparse': lambda args: f_sparse_op(args),
    'f_sparse_comp': lambda args: f_sparse_op_comp(args),
    'f_dense_last': lambda args: f_dense_op_last(args),
    'f_sparse_last': lambda args: f_sparse_op_last(args),
    'a_max': lambda args: a_max_op(args),
    'a_mean': lambda args: a_mean_op(args),
    'a_sum': lambda args: a_sum_op(args),
}

MIXED_OPS_sf = {
    'sf_TransE': lambda args: sf_TransE_op(args),
    'sf_DisMult': lambda args: sf_DisMult_op(args),
    'sf_ConvE': lambda args: sf_ConvE_op(args),
}

PRE_OPS = ['pre_mult', 'pre_sub', 'pre_add']
FIRST_OPS = ['f_zero', 'f_identity', 'f_dense_comp', 'f_sparse_comp', 'f_comp']
MIDDLE_OPS = ['a_max', 'a_sum', 'a_mean']
LAST_OPS = ['f_zero', 'f_identity', 'f_dense_last', 'f_sparse_last']
SF_OPS = ['sf_TransE', 'sf_DisMult']


def _with_stats(y, stats):
    y.mrg_stats = stats
    return y


# ------------------------------------------------------------------ composition (K1)
class _pre_op(nn.Module):
    comp = 0

    def forward(self, g, src_emb, hr):
        """reference: operations_lp.py:71-98"""
        return K.ComposeRows.apply(src_emb, hr, self.comp)


class pre_mult_op(_pre_op):
    comp = 1


class pre_sub_op(_pre_op):
    comp = 0


class pre_add_op(_pre_op):
    comp = 2


# ------------------------------------------------------------------ trivial filters
class f_identity_op(nn.Module):
    def forward(self, g, src_emb, src_emb_in):
        """reference: operations_lp.py:204-210"""
        return src_emb


class f_zero_op(nn.Module):
    def forward(self, g, src_emb, src_emb_in):
        """reference: operations_lp.py:214-220 (0 * src_emb keeps the autograd edge)"""
        return 0 * src_emb


# ------------------------------------------------------------------ aggregators (K5)
USE_TENSOR_CORES = True  # fused tcgen05 a_max kernel when the feature dim allows it (D % 8 == 0, D <= 256)


class a_max_op(nn.Module):
    """reference: operations_lp.py:223-235"""
    kind = 2

    def __init__(self, args):
        super().__init__()
        feature_dim = args.get('feature_dim', 100)
        self.linear = nn.Linear(feature_dim, feature_dim)

    def forward(self, block, src_emb, src_emb_in):
        E = block.num_edges()
        if self.kind == 2 and USE_TENSOR_CORES and K.amax_tc_supported(src_emb.shape[1]):
            return K.AMaxTC.apply(src_emb, self.linear.weight, self.linear.bias, block, True)
        m_pre = K.linear(self.linear, src_emb[:E, :])  # edge-tile GEMM (tcgen05, bias fused); ReLU applied on load by the reducer
        return K.SegReduce.apply(m_pre, src_emb[E:, :], block, self.kind, True)


class a_mean_op(a_max_op):
    """reference: operations_lp.py:238-250"""
    kind = 1


class a_sum_op(nn.Module):
    """reference: operations_lp.py:252-264 (Dropout applies to the aggregate only)"""

    def __init__(self, args):
        super().__init__()
        self.drop_aggr = args.get('drop_aggr', 0.1)
        self.drop_sum = nn.Dropout(self.drop_aggr)

    def forward(self, block, src_emb, src_emb_in):
        if self.training and self.drop_aggr > 0:
            E = block.num_edges()
            agg = K.SegReduce.apply(src_emb[:E, :], None, block, 0, False)
            return self.drop_sum(agg) + src_emb[E:, :]
        return K.AggSumLP.apply(src_emb, block)


# ------------------------------------------------------------------ sparse gates (K3/K6)
def _comp_bounds(g, rows):
    E = g.num_edges()
    half = getattr(g, 'half', None)
    half = E // 2 if half is None else half   # a destination partition holds unequal direction halves
    return [(0, half), (half, E), (E, rows)], E


class f_sparse_op_comp(nn.Module):
    """reference: operations_lp.py:304-343"""

    def __init__(self, args):
        super().__init__()
        self._feature_dim = args.get('feature_dim', 100)
        D = self._feature_dim
        self.W_in = nn.Linear(2 * D, D, bias=True)
        self.a_in = nn.Linear(D, 1, bias=False)
        self.W_out = nn.Linear(2 * D, D, bias=True)
        self.a_out = nn.Linear(D, 1, bias=False)
        self.W_self = nn.Linear(2 * D, D, bias=True)
        self.a_self = nn.Linear(D, 1, bias=False)

    def forward(self, g, src_emb, src_emb_in):
        D = self._feature_dim
        bounds, E = _comp_bounds(g, src_emb.shape[0])
        v1, v2, c = K.collapse_gates(D, ((self.W_in, self.a_in), (self.W_out, self.a_out), (self.W_self, self.a_self)))
        y, stats = K.SparseGate.apply(src_emb, src_emb_in, v1, v2, c, bounds, g.norm(), E, (1 / 3, 1 / 3, 1 / 3))
        return _with_stats(y, stats)


class f_sparse_op(nn.Module):
    """reference: operations_lp.py:345-354"""

    def __init__(self, args):
        super().__init__()
        self._feature_dim = args.get('feature_dim', 100)
        self.W = nn.Linear(2 * self._feature_dim, self._feature_dim, bias=True)
        self.a = nn.Linear(self._feature_dim, 1, bias=False)

    def forward(self, g, src_emb, src_emb_in):
        D = self._feature_dim
        v1, v2, c = K.collapse_gates(D, ((self.W, self.a),))
        y, stats = K.SparseGate.apply(src_emb, src_emb_in, v1, v2, c, [(0, src_emb.shape[0])], None, 0, (1.0,))
        return _with_stats(y, stats)


class f_sparse_op_last(nn.Module):
    """reference: operations_lp.py:405-416"""

    def __init__(self, args):
        super().__init__()
        self._feature_dim = args.get('feature_dim', 100)
        self.W = nn.Linear(self._feature_dim, self._feature_dim, bias=True)
        self.a = nn.Linear(self._feature_dim, 1, bias=False)

    def forward(self, g, src_emb, src_emb_in):
        v1, _, c = K.collapse_gates(self._feature_dim, ((self.W, self.a),))
        y, stats = K.SparseGate.apply(src_emb, None, v1, None, c, [(0, src_emb.shape[0])], None, 0, (1.0,))
        return _with_stats(y, stats)


# ------------------------------------------------------------------ dense gates (K4)
def _seg_linear(W, x, xin, lo, hi):
    """W [x, xin] (+b) on rows [lo,hi): the edge-tile GEMM of the dense candidates (operations_lp.py:275-283,
    366-384) on the tcgen05 main loop (3xTF32, K = 2D) when the tile is large enough, else the library GEMM."""
    return K.linear(W, torch.cat([x[lo:hi], xin[lo:hi]], 1))


class f_dense_op_comp(nn.Module):
    """reference: operations_lp.py:356-390"""

    def __init__(self, args):
        super().__init__()
        self._feature_dim = args.get('feature_dim', 100)
        D = self._feature_dim
        self.W_in = nn.Linear(2 * D, D, bias=True)
        self.W_out = nn.Linear(2 * D, D, bias=True)
        self.W_self = nn.Linear(2 * D, D, bias=True)

    def forward(self, g, src_emb, src_emb_in):
        bounds, E = _comp_bounds(g, src_emb.shape[0])
        z = torch.cat([_seg_linear(W, src_emb, src_emb_in, lo, hi)
                       for W, (lo, hi) in zip((self.W_in, self.W_out, self.W_self), bounds)], 0)
        y, stats = K.DenseGate.apply(z, src_emb, True, g.norm(), E, (1 / 3, 1 / 3, 1 / 3), bounds)
        return _with_stats(y, stats)


class f_comp_op(nn.Module):
    """reference: operations_lp.py:266-288 (bias-free; self rows are NOT scaled by 1/3)"""

    def __init__(self, args):
        super().__init__()
        self._feature_dim = args.get('feature_dim', 100)
        D = self._feature_dim
        self.W_in = nn.Linear(2 * D, D, bias=False)
        self.W_out = nn.Linear(2 * D, D, bias=False)
        self.W_self = nn.Linear(2 * D, D, bias=False)

    def forward(self, g, src_emb, src_emb_in):
        bounds, E = _comp_bounds(g, src_emb.shape[0])
        z = torch.cat([_seg_linear(W, src_emb, src_emb_in, lo, hi)
                       for W, (lo, hi) in zip((self.W_in, self.W_out, self.W_self), bounds)], 0)
        y, stats = K.DenseGate.apply(z, src_emb, False, g.norm(), E, (1 / 3, 1 / 3, 1.0), bounds)
        return _with_stats(y, stats)


class f_dense_op(nn.Module):
    """reference: operations_lp.py:290-301"""

    def __init__(self, args):
        super().__init__()
        self._feature_dim = args.get('feature_dim', 100)
        self.W = nn.Linear(2 * self._feature_dim, self._feature_dim, bias=True)

    def forward(self, g, src_emb, src_emb_in):
        rows = src_emb.shape[0]
        z = _seg_linear(self.W, src_emb, src_emb_in, 0, rows)
        y, stats = K.DenseGate.apply(z, src_emb, True, None, 0, (1.0,), [(0, rows)])
        return _with_stats(y, stats)


class f_dense_op_last(nn.Module):
    """reference: operations_lp.py:392-401"""

    def __init__(self, args):
        super().__init__()
        self._feature_dim = args.get('feature_dim', 100)
        self.W = nn.Linear(self._feature_dim, self._feature_dim, bias=True)

    def forward(self, g, src_emb, src_emb_in):
        rows = src_emb.shape[0]
        y, stats = K.DenseGate.apply(K.linear(self.W, src_emb), src_emb, True, None, 0, (1.0,), [(0, rows)])
        return _with_stats(y, stats)


# ------------------------------------------------------------------ score functions
class sf_DisMult_op(nn.Module):
    """reference: operations_lp.py:115-127.  forward returns probabilities [B, N] (predict()
    ranks with them); the training loss goes through ``loss`` = fused sigmoid+BCE kernel."""

    def __init__(self, args):
        super().__init__()

    def logits(self, all_ent, sub_emb, rel_emb):
        return torch.mm(sub_emb * rel_emb, all_ent.transpose(1, 0))

    def forward(self, all_ent, sub_emb, rel_emb):
        return torch.sigmoid(self.logits(all_ent, sub_emb, rel_emb))

    def loss(self, all_ent, sub_emb, rel_emb, label):
        if USE_TENSOR_CORES and all_ent.is_cuda and K.distmult_bce_supported(all_ent.shape[1]):
            return K.DistMultBCE.apply(all_ent, sub_emb, rel_emb, label)     # fused tcgen05 GEMM + BCE
        return K.SigmoidBCE.apply(self.logits(all_ent, sub_emb, rel_emb), label)


class sf_TransE_op(nn.Module):
    """reference: operations_lp.py:101-112 (selectable by genotype):
    sigmoid(gamma - ||sub_emb + rel_emb - all_ent||_1) as a fused L1-distance kernel (csrc/score.cu) instead of the
    reference's [B, N, D] broadcast; `loss` feeds the logits to the fused sigmoid+BCE kernel like sf_DisMult."""

    def __init__(self, args):
        super().__init__()
        self.gamma = args.get('gamma', 40)

    def logits(self, all_ent, sub_emb, rel_emb):
        if all_ent.is_cuda and all_ent.shape[1] % 4 == 0:
            return K.TransELogits.apply(all_ent, sub_emb + rel_emb, self.gamma)
        raise RuntimeError("sf_TransE runs on the CUDA path only (feature dim must be a multiple of 4)")

    def forward(self, all_ent, sub_emb, rel_emb):
        return torch.sigmoid(self.logits(all_ent, sub_emb, rel_emb))

    def loss(self, all_ent, sub_emb, rel_emb, label):
        return K.SigmoidBCE.apply(self.logits(all_ent, sub_emb, rel_emb), label)


class sf_ConvE_op(nn.Module):
    """reference: operations_lp.py:130-200 (CNN scorer, out of the MP hot path; plain torch)."""

    def __init__(self, args):
        super().__init__()
        self.embed_dim = args.get('embed_dim', 200)
        self.conve_hid_drop, self.feat_drop = args.get('conve_hid_drop', 0.3), args.get('feat_drop', 0.3)
        self.num_filt = args.get('num_filt', 200)
        self.ker_sz, self.k_w, self.k_h = args.get('ker_sz', 7), args.get('k_w', 10), args.get('k_h', 20)
        self.bn0 = nn.BatchNorm2d(1)
        self.bn1 = nn.BatchNorm2d(self.num_filt)
        self.bn2 = nn.BatchNorm1d(self.embed_dim)
        self.feature_drop = nn.Dropout(self.feat_drop)
        self.hidden_drop = nn.Dropout(self.conve_hid_drop)
        self.conv2d = nn.Conv2d(1, self.num_filt, (self.ker_sz, self.ker_sz), 1, 0, bias=True)
        self.flat_sz = (2 * self.k_h - self.ker_sz + 1) * (self.k_w - self.ker_sz + 1) * self.num_filt
        self.fc = nn.Linear(self.flat_sz, self.embed_dim)

    def forward(self, all_ent, sub_emb, rel_emb):
        assert self.embed_dim == self.k_h * self.k_w
        x = torch.cat([sub_emb.view(-1, 1, self.embed_dim), rel_emb.view(-1, 1, self.embed_dim)], 1)
        x = self.bn0(x.reshape(-1, 1, 2 * self.k_h, self.k_w))
        x = self.feature_drop(F.relu(self.bn1(self.conv2d(x))))
        x = self.hidden_drop(self.fc(x.view(-1, self.flat_sz)))
        x = F.relu(self.bn2(x))
        return torch.sigmoid(torch.mm(x, all_ent.transpose(1, 0)))
