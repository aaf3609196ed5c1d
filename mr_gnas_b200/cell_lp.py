"""LP supernet cells under the reference's names (models/cell_lp.py)."""
import torch.nn as nn

from . import supernet as S
from .operations_lp import FIRST_OPS, LAST_OPS, MIDDLE_OPS, MIXED_OPS, MIXED_OPS_sf, PRE_OPS, SF_OPS


class MixedOp(S.MixedOp):
    """reference: cell_lp.py:12-33"""

    def __init__(self, feature_dim, drop_aggr, operations):
        super().__init__(MIXED_OPS, feature_dim, operations, {'feature_dim': feature_dim, 'drop_aggr': drop_aggr},
                         with_linear=False)
        self._drop_aggr = drop_aggr


class Cell(S.SuperCell):
    """reference: cell_lp.py:155-188"""

    def __init__(self, nb_zero_nodes, nb_first_nodes, nb_last_nodes, feature_dim, dropout_aggr):
        super().__init__(MIXED_OPS, (PRE_OPS, FIRST_OPS, MIDDLE_OPS, LAST_OPS), nb_zero_nodes, nb_first_nodes,
                         nb_last_nodes, feature_dim, {'feature_dim': feature_dim, 'drop_aggr': dropout_aggr},
                         with_linear=False, nc_tail=False)


class MixedOp_SF(nn.Module):
    """reference: cell_lp.py:36-50 (constructed by the search network, never called: model_search_lp.py:160-161)."""

    def __init__(self, gamma, operations):
        super().__init__()
        self._ops = nn.ModuleList([nn.ModuleList([MIXED_OPS_sf[name]({'gamma': gamma})]) for name in operations])

    def forward(self, weights, all_ent, sub_emb, rel_emb):
        return sum(w * op[0](all_ent, sub_emb, rel_emb) for w, op in zip(weights, self._ops))


class Cell_Final(nn.Module):
    def __init__(self, gamma):
        super().__init__()
        self._ops = nn.ModuleList([MixedOp_SF(gamma, operations=SF_OPS)])

    def forward(self, all_ent, sub_emb, rel_emb, weights):
        return self._ops[0](weights[0], all_ent, sub_emb, rel_emb)


class Cell_SF(nn.Module):
    """reference: cell_lp.py:191-200"""

    def __init__(self, gamma):
        super().__init__()
        self.cell_score = Cell_Final(gamma)

    def forward(self, all_ent_emb, sub_emb, rel_emb, weights_sf):
        return self.cell_score(all_ent_emb, sub_emb, rel_emb, weights_sf)
