"""DARTS supernet cells shared by the LP and NC search networks (reference: models/cell_lp.py and
models/cell.py, which differ only in the candidate stack -- NC inserts Linear(D,D) before the BN --
and in the cell tail).  Module / attribute names follow the reference so supernet state_dicts match:
``cell_zero._ops.0._ops.<k>.<j>`` etc.

MixedOp.forward = sum_k w_k * ReLU(BN_k([Linear_k] op_k(g, h, h_in)))  (cell_lp.py:25-33, cell.py:23-31)
runs as ONE fused kernel over all candidates (mrg_mixed_sum_fwd) after each candidate's producer has
written its pre-BN output and column statistics."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as K
from .operations_lp import _pre_op

_pre_forward = _pre_op.forward


class MixedOp(nn.Module):
    def __init__(self, registry, feature_dim, operations, op_args, with_linear):
        super().__init__()
        self._feature_dim = feature_dim
        self._operations = operations
        self._with_linear = with_linear
        stacks = []
        for name in operations:
            mods = [registry[name](op_args)]
            if with_linear:
                mods.append(nn.Linear(feature_dim, feature_dim, bias=True))
            mods += [nn.BatchNorm1d(feature_dim), nn.ReLU()]
            stacks.append(nn.ModuleList(mods))
        self._ops = nn.ModuleList(stacks)

    def forward(self, weights, g, h, h_in):
        comps = [getattr(stack[0], 'comp', None) for stack in self._ops]
        if K.MIXED_PRE_FUSED and not self._with_linear and h.is_cuda and len(comps) <= 3 and h.shape == h_in.shape \
                and all(c is not None and type(stack[0]).forward is _pre_forward for c, stack in zip(comps, self._ops)):
            # every candidate is an elementwise composition of the SAME two rows: one shared read, no candidate outputs
            return K.mixed_pre(weights, h, h_in, comps, [stack[-2] for stack in self._ops])
        ys, bns = [], []
        for stack in self._ops:
            y = stack[0](g, h, h_in)
            if self._with_linear:
                y = K.linear(stack[1], y.float())
            ys.append(y)
            bns.append(stack[-2])
        return K.mixed_sum(weights, ys, bns)


def _sum(ts):
    ts = list(ts)
    out = ts[0]
    for t in ts[1:]:
        out = out + t
    return out


class CellZero(nn.Module):
    def __init__(self, make_mixed, pre_ops):
        super().__init__()
        self._ops = nn.ModuleList([make_mixed(pre_ops)])

    def forward(self, g, h, hr, weights):
        return self._ops[0](weights[0], g, h, hr)


class CellGrow(nn.Module):
    """Cell_First (in_nodes = 1 zero state) and Cell_Last (in_nodes = #middle outputs): node i sums one
    MixedOp per already available state (cell_lp.py:86-108, 130-152)."""

    def __init__(self, make_mixed, ops, in_nodes, nodes, keep_inputs):
        super().__init__()
        self._nodes, self._keep_inputs, self._in_nodes = nodes, keep_inputs, in_nodes
        self._ops = nn.ModuleList([make_mixed(ops) for i in range(nodes) for _ in range(i + in_nodes)])

    def forward(self, g, states, h_in, weights):
        states = list(states)
        offset = 0
        for _ in range(self._nodes):
            s = _sum(self._ops[offset + j](weights[offset + j], g, h, h_in) for j, h in enumerate(states))
            offset += len(states)
            states.append(s)
        return states if self._keep_inputs else states[self._in_nodes:]


class CellMiddle(nn.Module):
    def __init__(self, make_mixed, ops, nodes):
        super().__init__()
        self._nodes = nodes
        self._ops = nn.ModuleList([make_mixed(ops) for _ in range(nodes)])

    def forward(self, g, states, h_in, weights):
        return [self._ops[i](weights[i], g, states[i], h_in) for i in range(self._nodes)]


class SuperCell(nn.Module):
    """reference: cell_lp.py:155-188 (tail=None) and cell.py:118-146 (tail = BN -> ReLU -> dropout)."""

    def __init__(self, registry, op_lists, nb_zero_nodes, nb_first_nodes, nb_last_nodes, feature_dim, op_args,
                 with_linear, nc_tail, dropout=0.0):
        super().__init__()
        pre, first, middle, last = op_lists
        mk = lambda ops: MixedOp(registry, feature_dim, ops, op_args, with_linear)
        self._nb_zero_nodes, self._nb_first_nodes, self._nb_last_nodes = nb_zero_nodes, nb_first_nodes, nb_last_nodes
        self._feature_dim, self._dropout = feature_dim, dropout
        self.cell_zero = CellZero(mk, pre)
        self.cell_first = CellGrow(mk, first, 1, nb_first_nodes, keep_inputs=False)
        self.cell_middle = CellMiddle(mk, middle, nb_first_nodes)
        self.cell_last = CellGrow(mk, last, nb_first_nodes, nb_last_nodes, keep_inputs=True)
        self.concat_weights = nn.Linear((nb_first_nodes + nb_last_nodes) * feature_dim, feature_dim)
        self._nc_tail = nc_tail
        if nc_tail:
            self.batchnorm_h = nn.BatchNorm1d(feature_dim)
            self.activate = nn.ReLU()

    def forward(self, g, src_emb, hr, weights_zero, weights_first, weights_middle, weights_last):
        h_in = self.cell_zero(g, src_emb, hr, weights_zero)
        states = self.cell_first(g, [h_in], h_in, weights_first)
        states = self.cell_middle(g, states, h_in, weights_middle)
        states = self.cell_last(g, states, h_in, weights_last)
        h = self.concat_weights(torch.cat(states, dim=1))
        if self._nc_tail:
            h = K.bn_act(h, self.batchnorm_h, relu=True)
            h = F.dropout(h, self._dropout, training=self.training)
        return h


def softmax_rows(alphas, layer, n_edges):
    """softmax over the candidate axis of one layer's slice of an alpha table (model_search_lp.py:196-213)."""
    return F.softmax(alphas[layer * n_edges:(layer + 1) * n_edges], dim=1)


def decode_genotype(Genotype, W_zero, W_first, W_middle, W_last, op_lists, nb_zero_nodes, nb_first_nodes,
                    nb_last_nodes):
    """Discretise softmaxed alphas into a genotype (model_search_lp.py:215-311; model_search.py mirrors it):
    zero cell: argmax op per edge; first/last cells: for every new node pick the incoming edge whose best
    non-'f_zero' weight is largest, then that edge's best non-'f_zero' op; middle cell: argmax op."""
    pre_ops, first_ops, middle_ops, last_ops = op_lists
    gene = []
    prev = list(range(nb_zero_nodes))
    for n in range(nb_zero_nodes):
        gene.append((pre_ops[int(torch.argmax(W_zero[n]))], n + 1, prev[n]))
        prev[n] = n + 1

    def best_edge(W, n_in, ops):
        skip = ops.index('f_zero')
        cols = [k for k in range(len(ops)) if k != skip]
        j = sorted(range(n_in), key=lambda x: -max(W[x][k] for k in cols))[0]
        k_best = None
        for k in cols:
            if k_best is None or W[j][k] > W[j][k_best]:
                k_best = k
        return j, k_best

    start, base = 0, max(prev)
    for n in range(1, nb_first_nodes + 1):
        j, k = best_edge(W_first[start:start + n], n, first_ops)
        gene.append((first_ops[k], base + n, base + j))
        start += n
    concat = []
    mids = list(range(2, 2 + nb_first_nodes))
    for n in range(nb_first_nodes):
        new = max(mids) + 1
        gene.append((middle_ops[int(torch.argmax(W_middle[n]))], new, mids[n]))
        concat.append(new)
        mids[n] = new
    start = 0
    for n in range(nb_last_nodes):
        node_id = n + max(mids) + 1
        n_in = nb_first_nodes + n
        j, k = best_edge(W_last[start:start + n_in], n_in, last_ops)
        pre = mids[j] if j < nb_first_nodes else j - nb_first_nodes + max(mids) + 1
        gene.append((last_ops[k], node_id, pre))
        concat.append(node_id)
        start += n_in
    return Genotype(alpha_cell=gene, concat_node=concat, score_func=None)
