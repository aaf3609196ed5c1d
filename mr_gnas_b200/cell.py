"""NC supernet cells under the reference's names (models/cell.py)."""
from . import supernet as S
from .operations import FIRST_OPS, LAST_OPS, MIDDLE_OPS, MIXED_OPS, PRE_OPS


class MixedOp(S.MixedOp):
    """reference: cell.py:11-31 (candidate = op -> Linear(D,D) -> BN -> ReLU)"""

    def __init__(self, feature_dim, operations):
        super().__init__(MIXED_OPS, feature_dim, operations, {'feature_dim': feature_dim}, with_linear=True)


class Cell(S.SuperCell):
    """reference: cell.py:118-146"""

    def __init__(self, nb_zero_nodes, nb_first_nodes, nb_last_nodes, feature_dim, dropout=0.0):
        super().__init__(MIXED_OPS, (PRE_OPS, FIRST_OPS, MIDDLE_OPS, LAST_OPS), nb_zero_nodes, nb_first_nodes,
                         nb_last_nodes, feature_dim, {'feature_dim': feature_dim}, with_linear=True, nc_tail=True,
                         dropout=dropout)
