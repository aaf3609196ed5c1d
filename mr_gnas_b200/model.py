"""NC derived network (reference: models/model.py) on libmrgnas.  Blocks are MRBlock objects
(mr_gnas_b200.graph) instead of DGL blocks; the per-destination Python remap loop of
model.py:176-178 (O(N_dst * E) host work) is one device-side table lookup."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dist as D_
from . import functional as K
from .operations import MIXED_OPS


class OpModule(nn.Module):
    """reference: model.py:10-28 (op -> Linear(D,D) -> [BN if args.op_norm] -> ReLU)"""

    def __init__(self, args, operation_name):
        super().__init__()
        self.args = args
        self._feature_dim = args.feature_dim
        self.op = MIXED_OPS[operation_name]({'feature_dim': self._feature_dim})
        self.linear = nn.Linear(self._feature_dim, self._feature_dim, bias=True)
        self.batchnorm_h = nn.BatchNorm1d(self._feature_dim)
        self.activate = nn.ReLU()

    def forward(self, g, h, h_in):
        h = K.linear(self.linear, self.op(g, h, h_in))     # tcgen05 (3xTF32) on full-graph / large blocks
        if self.args.op_norm:
            return K.bn_act(h, self.batchnorm_h, relu=True)
        return K.ReluAct.apply(h)


class Cell(nn.Module):
    """reference: model.py:31-67"""

    def __init__(self, args, genotype):
        super().__init__()
        self.args = args
        self._genotype = genotype
        self._nb_nodes = len(set(edge[1] for edge in genotype.alpha_cell))
        self._feature_dim = args.feature_dim
        self._concat_node = list(range(1, 1 + self._nb_nodes)) if genotype.concat_node is None else genotype.concat_node
        self.batchnorm_h = nn.BatchNorm1d(self._feature_dim)
        self.activate = nn.ReLU()
        self._ops = nn.ModuleList([nn.ModuleList([nn.ModuleList() for _ in range(n)])
                                   for n in range(1, 1 + self._nb_nodes)])
        for (op_name, center_node, pre_node) in genotype.alpha_cell:
            self._ops[center_node - 1][pre_node].append(OpModule(args, op_name))
        self.concat = nn.Linear(len(self._concat_node) * self._feature_dim, self._feature_dim)

    def forward(self, g, src_emb, hr):
        zero_out = self._ops[0][0][0](g, src_emb, hr)
        states = [src_emb, zero_out]
        for n in range(1, self._nb_nodes):
            hs = [self._ops[n][i][0](g, states[i], zero_out) for i in range(n + 1) if len(self._ops[n][i]) > 0]
            states.append(hs[0] if len(hs) == 1 else sum(hs))
        h = K.linear(self.concat, torch.cat([states[idx] for idx in self._concat_node], dim=1))
        return K.bn_act(h, self.batchnorm_h, relu=True)


class MLPClassifier(nn.Module):
    """reference: model.py:70-85 (plain torch; not on the MP path)"""

    def __init__(self, input_dim, output_dim, L=2):
        super().__init__()
        layers = [nn.Linear(input_dim // 2 ** l, input_dim // 2 ** (l + 1), bias=True) for l in range(L)]
        layers.append(nn.Linear(input_dim // 2 ** L, output_dim, bias=True))
        self.FC_layers = nn.ModuleList(layers)
        self.L = L

    def forward(self, x):
        for l in range(self.L):
            x = F.relu(self.FC_layers[l](x))
        return self.FC_layers[self.L](x)


class mean_aggre(nn.Module):
    """reference: model.py:93-104 (registered but unused by forward)"""

    def __init__(self, feature_dim):
        super().__init__()
        self.linear = nn.Linear(feature_dim, feature_dim)

    def forward(self, block, src_emb):
        return K.SegReduce.apply(K.linear(self.linear, src_emb), None, block, 1, True)


def block_inputs(trip_index, blocks):
    """Per block: global source ids and edge types of its edges (model.py:153-165)."""
    src, et = [], []
    for b in blocks:
        sel = torch.index_select(trip_index, 0, b.edata['_ID'])
        src.append(sel[:, 1])
        et.append(b.edata['_TYPE'])
    return src, et


def remap_sources(next_src, dst_nid, num_nodes):
    """position of every next-block source inside this block's destination list (model.py:175-179)."""
    table = torch.full((num_nodes,), -1, dtype=torch.long, device=dst_nid.device)
    table[dst_nid] = torch.arange(dst_nid.numel(), device=dst_nid.device)
    return table[next_src]


def edge_type_features(linear, rel_table, etype, block):
    """`embedding_e_init(rel_table[etype])` (model.py:168-170).  The Linear is row-wise, so it is applied to the few rows of
    the relation table and THEN gathered per edge -- E_b x init_dim fewer GEMM rows forward and backward -- and the gather's
    backward is the deterministic segmented sum over the block's edge-type segments instead of ATen's sort-based index_put,
    which serialises on the duplicates (millions of edges over a few hundred relations)."""
    if not (K.NC_REL_REORDER and rel_table.is_cuda and etype.numel() > 0):
        return linear(rel_table[etype])
    feat = linear(rel_table)
    return K.gather_rows(feat, etype, K.key_segments(etype, feat.shape[0], block, 'etype'))


class Network(nn.Module):
    """reference: model.py:107-199"""

    def __init__(self, device, genotype, number_of_nodes, num_classes, num_rels, layers, zero_nodes, nodes,
                 feature_dim, init_fea_dim, num_base_r, criterion, args):
        super().__init__()
        self._device = device
        self._layers = layers
        self._in_dim_n, self._in_dim_e = number_of_nodes, num_rels
        self._feature_dim, self._init_fea_dim = feature_dim, init_fea_dim
        self._num_base_r, self._num_classes = num_base_r, num_classes
        self._criterion = criterion
        self.embedding_h = nn.Embedding(self._in_dim_n, self._init_fea_dim)
        self.embedding_e = nn.Embedding(self._num_base_r, self._init_fea_dim)
        self.rel_wt = self.get_param([self._in_dim_e, self._num_base_r])
        self.embedding_h_init = nn.Linear(self._init_fea_dim, self._feature_dim, bias=False)
        self.embedding_e_init = nn.Linear(self._init_fea_dim, self._feature_dim, bias=False)
        self.cells = nn.ModuleList([Cell(args, genotype[i]) for i in range(self._layers)])
        self.classifier = MLPClassifier(self._feature_dim, self._num_classes)
        self.mean_aggre = mean_aggre(self._feature_dim)
        self.batchnorm_h = nn.BatchNorm1d(self._feature_dim)
        self.activate = nn.ReLU()

    def get_param(self, shape):
        param = nn.Parameter(torch.Tensor(*shape))
        nn.init.xavier_normal_(param, gain=nn.init.calculate_gain('relu'))
        return param

    def _cell(self, i, cell, block, src_embed, edges_embed):
        return cell(block, src_embed, edges_embed)

    def _forward_partitioned(self, trip_index, block, part):
        """Full-graph layers on ONE destination range (dist.nc_partition): every layer reads the
        sources of its edges from the full node table (layer 0: the replicated input embedding; later layers: the
        owned rows of every rank all-gathered over NVLink), BatchNorm statistics are summed over the ranks.
        Returns this rank's rows [hi - lo, D] of the final node embedding."""
        src_ls, et_ls = block_inputs(trip_index, block)
        rel_table = torch.mm(self.rel_wt, self.embedding_e.weight)
        node_embed = None
        with D_.use(part):
            for i, cell in enumerate(self.cells):
                if i == 0:
                    src_embed = self.embedding_h_init(self.embedding_h(src_ls[i]))
                else:
                    src_embed = D_.AllGatherRows.apply(node_embed, part)[src_ls[i]]
                edges_embed = edge_type_features(self.embedding_e_init, rel_table, et_ls[i], block[i])
                node_embed = self._cell(i, cell, block[i], src_embed, edges_embed)
            return K.bn_act(node_embed, self.batchnorm_h, relu=True)

    def _loss_partitioned(self, trip_index, block, labels, idx):
        """Cross-entropy over the labelled nodes `idx` (global ids, labels[idx] their classes) of a partitioned
        full-graph forward: every rank classifies the labelled nodes it owns; the mean is taken over ALL of them."""
        part = block[0].part
        logits = self.classifier(self._forward_partitioned(trip_index, block, part))
        own = (idx >= part.lo) & (idx < part.hi)
        loc = (idx - part.lo).clamp(0, max(part.n_local - 1, 0))
        per = F.cross_entropy(logits[loc], labels[idx], reduction='none') * own.to(logits.dtype)
        return D_.AllReduceSum.apply(per.sum() / idx.numel(), part)

    def _forward(self, trip_index, block):
        if getattr(block[0], 'part', None) is not None:
            return self._forward_partitioned(trip_index, block, block[0].part)
        src_ls, et_ls = block_inputs(trip_index, block)
        rel_table = torch.mm(self.rel_wt, self.embedding_e.weight)  # rows gathered per edge below
        node_embed = src_embed = None
        for i, cell in enumerate(self.cells):
            if i == 0:
                src_embed = self.embedding_h_init(self.embedding_h(src_ls[i]))
            edges_embed = edge_type_features(self.embedding_e_init, rel_table, et_ls[i], block[i])
            node_embed = self._cell(i, cell, block[i], src_embed, edges_embed)
            if i < len(src_ls) - 1:
                src_embed = node_embed[remap_sources(src_ls[i + 1], block[i].dstdata['_ID'], self._in_dim_n)]
        return K.bn_act(node_embed, self.batchnorm_h, relu=True)

    def forward(self, trip_index, g):
        return self.classifier(self._forward(trip_index, g))

    def _loss(self, trip_index, g, labels, idx):
        return self._criterion(self.forward(trip_index, g), labels[idx])
