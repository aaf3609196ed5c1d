"""NC DARTS supernet (reference: models/model_search.py): same constructor, parameter registration and
alpha-table layout; cells are mr_gnas_b200.cell.Cell (fused MixedOp kernels), blocks are MRBlock."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as K
from .cell import Cell
from .genotypes import Genotype
from .model import MLPClassifier, block_inputs, mean_aggre, remap_sources
from .operations import FIRST_OPS, LAST_OPS, MIDDLE_OPS, PRE_OPS
from .supernet import decode_genotype, softmax_rows


class Network(nn.Module):
    def __init__(self, device, number_of_nodes, num_classes, num_rels, layers, zero_nodes, nodes, feature_dim,
                 init_fea_dim, num_base_r, dropout=0.0):
        super().__init__()
        self._device, self._layers = device, layers
        self._in_dim_n, self._in_dim_e = number_of_nodes, num_rels
        self._feature_dim, self._init_fea_dim = feature_dim, init_fea_dim
        self._num_base_r, self._num_classes = num_base_r, num_classes
        self._criterion = nn.CrossEntropyLoss()
        self._nb_zero_nodes, self._nb_first_nodes, self._nb_last_nodes = zero_nodes, nodes, nodes
        self._nb_zero_edges = zero_nodes
        self._nb_first_edges = sum(zero_nodes + i for i in range(nodes))
        self._nb_middle_edges = nodes
        self._nb_last_edges = sum(nodes + i for i in range(nodes))
        self.embedding_h = nn.Embedding(number_of_nodes, init_fea_dim)
        self.embedding_e = nn.Embedding(num_base_r, init_fea_dim)
        self.rel_wt = self.get_param([num_rels, num_base_r])
        self.embedding_h_init = nn.Linear(init_fea_dim, feature_dim, bias=False)
        self.embedding_e_init = nn.Linear(init_fea_dim, feature_dim, bias=False)
        self.cells = nn.ModuleList([Cell(zero_nodes, nodes, nodes, feature_dim) for _ in range(layers)])
        self._initialize_alphas()
        self.classifier = MLPClassifier(feature_dim, num_classes)
        self.mean_aggre = mean_aggre(feature_dim)
        self.batchnorm_h = nn.BatchNorm1d(feature_dim)
        self.activate = nn.ReLU()
        self._dropout = dropout

    def get_param(self, shape):
        param = nn.Parameter(torch.Tensor(*shape))
        nn.init.xavier_normal_(param, gain=nn.init.calculate_gain('relu'))
        return param

    def _initialize_alphas(self):
        """reference: model_search.py:107-141"""
        L = self._layers
        mk = lambda rows, ops: (1e-3 * torch.randn(rows, len(ops))).to(self._device).requires_grad_(True)
        self.alphas_zero_cell = mk(self._nb_zero_edges * L, PRE_OPS)
        self.alphas_first_cell = mk(self._nb_first_edges * L, FIRST_OPS)
        self.alphas_middle_cell = mk(self._nb_middle_edges * L, MIDDLE_OPS)
        self.alphas_last_cell = mk(self._nb_last_edges * L, LAST_OPS)
        self._arch_parameters = [self.alphas_zero_cell, self.alphas_first_cell, self.alphas_middle_cell,
                                 self.alphas_last_cell]

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
        g = decode_genotype(Genotype, *W, (PRE_OPS, FIRST_OPS, MIDDLE_OPS, LAST_OPS), self._nb_zero_nodes,
                            self._nb_first_nodes, self._nb_last_nodes)
        return g

    def show_genotypes(self):
        return [self.show_genotype(i) for i in range(self._layers)]

    def _forward(self, trip_index, block):
        """reference: model_search.py:143-177"""
        src_ls, et_ls = block_inputs(trip_index, block)
        rel_table = torch.mm(self.rel_wt, self.embedding_e.weight)
        node_embed = src_embed = None
        for i, cell in enumerate(self.cells):
            if i == 0:
                src_embed = self.embedding_h_init(self.embedding_h(src_ls[i]))
            edges_embed = self.embedding_e_init(rel_table[et_ls[i]])
            node_embed = cell(block[i], src_embed, edges_embed, *self.show_weights(i))
            if i < len(src_ls) - 1:
                src_embed = node_embed[remap_sources(src_ls[i + 1], block[i].dstdata['_ID'], self._in_dim_n)]
        h = K.bn_act(node_embed, self.batchnorm_h, relu=True)
        return F.dropout(h, self._dropout, training=self.training)

    def forward(self, trip_index, g):
        return self.classifier(self._forward(trip_index, g))

    def _loss(self, trip_index, g, labels, idx):
        return self._criterion(self.forward(trip_index, g), labels[idx])
