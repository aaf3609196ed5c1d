"""Node-classification fine-grained operators: registry names, constructor dicts, parameter
names and ``forward(g, src_emb, src_emb_in)`` signatures of the reference's models/operations.py.

Differences from the LP set (operations_lp.py): aggregators reduce ALL E_b edge rows of a block
to its n_dst destination nodes (no self-loop rows, no residual) and there is an extra ``a_std``;
filters have no direction split and no degree norm.  The reference runs these through DGL's
Python UDF reducers (degree bucketing, operations.py:105-190); here they are the same segmented
reduction kernels as the LP path (mailbox order == ascending edge id)."""
import torch
import torch.nn as nn

from . import functional as K
from . import operations_lp as lp
from .operations_lp import (f_dense_op, f_dense_op_last, f_identity_op, f_sparse_op, f_sparse_op_last,  # noqa: F401
                            f_zero_op, pre_add_op, pre_mult_op, pre_sub_op)

MIXED_OPS = {
    'pre_mult': lambda args: pre_mult_op(),
    'pre_sub': lambda args: pre_sub_op(),
    'pre_add': lambda args: pre_add_op(),
    'f_zero': lambda args: f_zero_op(),
    'f_identity': lambda args: f_identity_op(),
    'f_dense': lambda args: f_dense_op(args),
    'f_sparse': lambda args: f_sparse_op(args),
    'f_dense_last': lambda args: f_dense_op_last(args),
    'f_sparse_last': lambda args: f_sparse_op_last(args),
    'a_max': lambda args: a_max_op(args),
    'a_mean': lambda args: a_mean_op(args),
    'a_sum': lambda args: a_sum_op(args),
    'a_std': lambda args: a_std_op(args),
}
PRE_OPS = ['pre_mult', 'pre_sub', 'pre_add']
FIRST_OPS = ['f_zero', 'f_identity', 'f_dense', 'f_sparse']
MIDDLE_OPS = ['a_max', 'a_sum', 'a_mean']
LAST_OPS = ['f_zero', 'f_identity', 'f_dense_last', 'f_sparse_last']
EPS = 1e-5


class a_max_op(nn.Module):
    """reference: operations.py:109-121"""
    kind = 2

    def __init__(self, args):
        super().__init__()
        feature_dim = args.get('feature_dim', 100)
        self.linear = nn.Linear(feature_dim, feature_dim)

    def forward(self, block, src_emb, src_emb_in):
        if self.kind == 2 and lp.USE_TENSOR_CORES and K.amax_tc_supported(src_emb.shape[1]):
            return K.AMaxTC.apply(src_emb, self.linear.weight, self.linear.bias, block, False)
        return K.SegReduce.apply(K.linear(self.linear, src_emb), None, block, self.kind, True)


class a_mean_op(a_max_op):
    """reference: operations.py:128-146"""
    kind = 1


class a_sum_op(nn.Module):
    """reference: operations.py:153-164"""

    def __init__(self, args):
        super().__init__()

    def forward(self, block, src_emb, src_emb_in):
        return K.SegReduce.apply(src_emb, None, block, 0, False)


class a_std_op(nn.Module):
    """reference: operations.py:167-190: sqrt(relu(E[h^2] - E[h]^2) + EPS) per destination; nodes without
    in-edges keep 0 (DGL never calls the UDF for them).  Two segmented MEAN reductions + node-level math."""

    def __init__(self, args):
        super().__init__()

    def forward(self, g, src_emb, src_emb_in):
        mean = K.SegReduce.apply(src_emb, None, g, 1, False)
        msq = K.SegReduce.apply(src_emb * src_emb, None, g, 1, False)
        out = torch.sqrt(torch.relu(msq - mean * mean) + EPS)
        return torch.where((g.in_deg > 0).view(-1, 1), out, torch.zeros_like(out))
