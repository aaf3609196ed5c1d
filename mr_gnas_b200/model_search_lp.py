"""LP DARTS supernet: constructor signature, parameter names / registration (and RNG draw) order,
alpha tables and genotype decoding of the reference's models/model_search_lp.py, with every
MixedOp evaluated by the fused libmrgnas kernels."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as K
from .cell_lp import Cell, Cell_SF
from .genotypes import Genotype
from .operations_lp import FIRST_OPS, LAST_OPS, MIDDLE_OPS, PRE_OPS, SF_OPS
from .supernet import decode_genotype, softmax_rows


class Network(nn.Module):
    def __init__(self, device, number_of_nodes, num_rels, layers, zero_nodes, first_nodes, last_nodes, feature_dim,
                 init_fea_dim, num_base_r, gamma, dropout_cell, drop_aggr):
        super().__init__()
        self._device = device
        self._layers = layers
        self._num_ent = number_of_nodes
        self._num_rel = num_rels * 2 + 1
        self._feature_dim = feature_dim
        self.num_base_r = num_base_r
        self.init_fea_dim = init_fea_dim
        self._nb_zero_nodes, self._nb_first_nodes, self._nb_last_nodes = zero_nodes, first_nodes, last_nodes
        self._nb_zero_edges = zero_nodes
        self._nb_final_nodes = self._nb_final_edges = 1
        self._nb_first_edges = sum(zero_nodes + i for i in range(first_nodes))
        self._nb_middle_edges = first_nodes
        self._nb_last_edges = sum(first_nodes + i for i in range(last_nodes))
        # construction order == reference (model_search_lp.py:41-79): it fixes the seeded init
        self.embedding_h = nn.Embedding(self._num_ent, self.init_fea_dim)
        self.embedding_e = nn.Embedding(self.num_base_r, self._feature_dim)
        self.linear_e = nn.Linear(self.init_fea_dim, self._feature_dim)
        self.rel_wt = self.get_param([self._num_rel, self.num_base_r])
        self.w_rel = self.get_param([self._feature_dim, self._feature_dim])
        self._drop_aggr = drop_aggr
        self.cells = nn.ModuleList([Cell(zero_nodes, first_nodes, last_nodes, feature_dim, drop_aggr)
                                    for _ in range(layers)])
        self._initialize_alphas()
        self.gamma = gamma
        self.score_func = Cell_SF(gamma)
        self.batchnorm_h = nn.BatchNorm1d(feature_dim)
        self.activate = nn.ReLU()
        self._dropout = dropout_cell

    def get_param(self, shape):
        param = nn.Parameter(torch.Tensor(*shape))
        nn.init.xavier_normal_(param, gain=nn.init.calculate_gain('relu'))
        return param

    # ---------------------------------------------------------------- architecture parameters
    def _initialize_alphas(self):
        """reference: model_search_lp.py:99-129 (same shapes, same randn order)."""
        L = self._layers
        mk = lambda rows, ops: (1e-3 * torch.randn(rows, len(ops))).to(self._device).requires_grad_(True)
        self.alphas_zero_cell = mk(self._nb_zero_edges * L, PRE_OPS)
        self.alphas_first_cell = mk(self._nb_first_edges * L, FIRST_OPS)
        self.alphas_middle_cell = mk(self._nb_middle_edges * L, MIDDLE_OPS)
        self.alphas_last_cell = mk(self._nb_last_edges * L, LAST_OPS)
        self.alphas_final_cell = mk(self._nb_final_edges, SF_OPS)
        self._arch_parameters = [self.alphas_zero_cell, self.alphas_first_cell, self.alphas_middle_cell,
                                 self.alphas_last_cell, self.alphas_final_cell]

    def arch_parameters(self):
        return self._arch_parameters

    def load_alpha(self, alphas):
        for x, y in zip(self.arch_parameters(), alphas):
            x.data.copy_(y.data)

    def show_weights(self, nb_layer):
        return (softmax_rows(self.alphas_zero_cell, nb_layer, self._nb_zero_edges),
                softmax_rows(self.alphas_first_cell, nb_layer, self._nb_first_edges),
                softmax_rows(self.alphas_middle_cell, nb_layer, self._nb_middle_edges),
                softmax_rows(self.alphas_last_cell, nb_layer, self._nb_last_edges))

    def show_genotype(self, nb_layer):
        W = [w.detach().cpu() for w in self.show_weights(nb_layer)]
        return decode_genotype(Genotype, *W, (PRE_OPS, FIRST_OPS, MIDDLE_OPS, LAST_OPS), self._nb_zero_nodes,
                               self._nb_first_nodes, self._nb_last_nodes)

    def show_genotypes(self):
        return [self.show_genotype(i) for i in range(self._layers)]

    # ---------------------------------------------------------------- forward / loss
    def _forward_lp(self, g_train, node_id, src_in, edge_type):
        """reference: model_search_lp.py:131-163."""
        dev = self.embedding_h.weight.device
        all_ent_emb = self.linear_e(self.embedding_h.weight)
        rel_embed = K.matmul(self.rel_wt, self.embedding_e.weight)
        nodes = g_train.nodes().to(dev)
        src_in_final = torch.cat((src_in, nodes), dim=0)
        src_id_final = node_id[src_in_final].reshape(-1)
        edge_self = torch.full((nodes.numel(),), self._num_rel - 1, dtype=torch.long, device=dev)
        edge_type_final = torch.cat((edge_type.long(), edge_self), dim=0)
        ent_emb = None
        # The gathers below read the graph's own edge-expanded index arrays (rows [0, E) = the edges in the graph's order,
        # rows [E, E + N) = the self loops), so their backward can be the graph's deterministic segmented sums instead of
        # ATen's index_put (K.GatherRows).  That holds when src_in / edge_type ARE the graph's arrays -- what
        # utils_rgcn hands to the search scripts -- and is verified once per graph object (one device comparison).
        own = getattr(g_train, 'csc', None) is not None and g_train.M == src_in_final.numel() \
            and g_train.n_rel_rows == rel_embed.shape[0] and node_id.numel() == g_train.N
        if own and not getattr(g_train, '_gather_verified', False):
            own = bool(torch.equal(src_in.to(torch.int32), g_train.src) & torch.equal(edge_type.to(torch.int32), g_train.etype))
            g_train._gather_verified = own
        for i, cell in enumerate(self.cells):
            W_zero, W_first, W_middle, W_last = self.show_weights(i)
            if own:
                nodes_emb = all_ent_emb[node_id.reshape(-1)] if i == 0 else ent_emb      # unique ids: a plain row gather
                ent_emb_in = K.gather_rows(nodes_emb, g_train.src_final, g_train.csc)
                hr = K.gather_rows(rel_embed, g_train.et_final, g_train.rel)
            else:
                ent_emb_in = all_ent_emb[src_id_final] if i == 0 else torch.cat((ent_emb[src_in], ent_emb), dim=0)
                hr = rel_embed[edge_type_final]
            ent_emb = cell(g_train, ent_emb_in, hr, W_zero, W_first, W_middle, W_last)
            relu = not (i == 0 and len(self.cells) != 1)  # layer 0 of a deeper net is not activated (:146-148)
            ent_emb = K.bn_act(ent_emb, self.batchnorm_h, relu=relu)
            ent_emb = F.dropout(ent_emb, self._dropout, training=self.training)
            rel_embed = K.matmul(rel_embed, self.w_rel)
        return ent_emb, rel_embed

    def forward(self, g_train, node_id, src_in, edge_type):
        return self._forward_lp(g_train, node_id, src_in, edge_type)

    def calc_score(self, ent_embedding, rel_embedding, triplets):
        """triplet-wise DistMult sum_d s*r*o (model_search_lp.py:169-176)"""
        s = ent_embedding[triplets[:, 0]]
        r = K.gather_few(rel_embedding, triplets[:, 1])
        o = ent_embedding[triplets[:, 2]]
        return torch.sum(s * r * o, dim=1)

    def get_loss(self, g_train, ent_embed, rel_embed, triplets, labels):
        return F.binary_cross_entropy_with_logits(self.calc_score(ent_embed, rel_embed, triplets), labels)

    def _loss(self, g_train, node_id, src_in, edge_type, triplets, labels):
        ent, rel = self.forward(g_train, node_id, src_in, edge_type)
        return F.binary_cross_entropy_with_logits(self.calc_score(ent, rel, triplets), labels)
