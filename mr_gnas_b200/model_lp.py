"""Genotype-compiled LP network: same classes, constructor arguments, parameter names and
registration order as the reference's models/model_lp.py (so ``model.apply(weights_init)``
draws the same seeded init and reference checkpoints load by key), running on libmrgnas.

Differences that are not visible through the module API:
  * BN+ReLU after each op is one fused pass fed by the producer kernel's column statistics;
  * Network._forward_lp never materialises all_ent_emb[src_id_final] / rel_embed[edge_type_final]
    (2 x [E+N, D]): the gather is fused into the pre_* composition kernel (model_lp.py:126-131);
  * ``_loss`` uses the fused sigmoid+BCE kernel instead of materialising probabilities.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dist as D_
from . import functional as K
from . import fused_cell
from .operations_lp import MIXED_OPS, MIXED_OPS_sf

USE_FUSED_CELL = True  # one autograd node for the edge-level chain of a cell (fused_cell.py) when the genotype allows


class OpModule(nn.Module):
    """reference: model_lp.py:13-35"""

    def __init__(self, args, operation_name):
        super().__init__()
        self.args = args
        self._feature_dim = args.feature_dim
        op_args = {'feature_dim': self._feature_dim, 'drop_aggr': args.drop_aggr}
        self.op = MIXED_OPS[operation_name](op_args)
        self.op_name = operation_name
        self.batchnorm_h = nn.BatchNorm1d(self._feature_dim)
        self.activate = nn.ReLU()
        self.drop_op = self.args.drop_op

    def _post(self, h):
        # model_lp.py:31: the guard is truthy for every op except 'pre_mult'; the dropout on :34
        # discards its result, so nothing is dropped.
        if self.op_name != 'pre_mult':
            h = K.bn_act(h, self.batchnorm_h, relu=True, stats=getattr(h, 'mrg_stats', None))
        return h

    def forward(self, g, h, h_in):
        return self._post(self.op(g, h, h_in))

    def forward_gathered(self, g, ent, rel):
        """pre_* op applied to the virtual gather ent[src_final] (.) rel[et_final]."""
        y, stats = K.GatherCompose.apply(ent, rel, g, self.op.comp)
        y.mrg_stats = stats
        return self._post(y)


class Cell(nn.Module):
    """reference: model_lp.py:38-74"""

    def __init__(self, args, genotype):
        super().__init__()
        self.args = args
        self._genotype = genotype
        self._nb_nodes = len(set([edge[1] for edge in genotype.alpha_cell]))
        self._feature_dim = args.feature_dim
        self._concat_node = list(range(1, 1 + self._nb_nodes)) if genotype.concat_node is None else genotype.concat_node
        self.batchnorm_h = nn.BatchNorm1d(self._feature_dim)
        self.activate = nn.ReLU()
        self._compile()

    def _compile(self):
        nb_nodes = self._nb_nodes
        self._ops = nn.ModuleList([nn.ModuleList([nn.ModuleList() for i in range(n)]) for n in range(1, 1 + nb_nodes)])
        for (op_name, center_node, pre_node) in self._genotype.alpha_cell:
            center_node -= 1
            self._ops[center_node][pre_node].append(OpModule(self.args, op_name))
        self.concat = nn.Linear(len(self._concat_node) * self._feature_dim, self._feature_dim)

    def uses_state0(self):
        """True if anything other than the pre_* op reads the raw gathered input."""
        return (0 in self._concat_node) or any(len(self._ops[n][0]) > 0 for n in range(1, self._nb_nodes))

    def _rest(self, g, states, zero_out):
        for n in range(1, self._nb_nodes):
            hs = []
            for i in range(n + 1):
                if len(self._ops[n][i]) > 0:
                    hs.append(self._ops[n][i][0](g, states[i], zero_out))
            states.append(hs[0] if len(hs) == 1 else sum(hs))  # 0 + t == t exactly; skip the extra pass
        h = K.linear(self.concat, torch.cat([states[idx] for idx in self._concat_node], dim=1))
        return K.bn_act(h, self.batchnorm_h, relu=True)

    def forward(self, g, src_emb, hr):
        zero_out = self._ops[0][0][0](g, src_emb, hr)
        return self._rest(g, [src_emb, zero_out], zero_out)

    def forward_fused(self, g, ent, rel):
        """Same result as forward(g, ent[src_final], rel[et_final]) without materialising either."""
        first = self._ops[0][0][0]
        plan = fused_cell.plan_for(self) if USE_FUSED_CELL else None
        if plan is not None:
            return self._forward_plan(g, ent, rel, plan)
        if not hasattr(first.op, 'comp') or self.uses_state0():
            src = ent[g.src_final.long()]
            return self.forward(g, src, rel[g.et_final.long()])
        zero_out = first.forward_gathered(g, ent, rel)
        return self._rest(g, [None, zero_out], zero_out)

    def _forward_plan(self, g, ent, rel, plan):
        """Edge-level chain in one fused autograd node; node-level ops through their modules."""
        _, gates, aggs = plan
        agg_out = fused_cell.run(self, g, ent, rel, plan)
        agg_mod = {node: om for node, om, _ in aggs}
        edge_nodes = {1} | {node for node, _, _ in gates}
        states = {}
        for n in range(1, self._nb_nodes):
            node = n + 1
            if node in edge_nodes:
                continue
            if node in agg_out:
                states[node] = agg_mod[node]._post(agg_out[node])
                continue
            hs = [self._ops[n][i][0](g, states[i], None) for i in range(n + 1) if len(self._ops[n][i]) > 0]
            states[node] = hs[0] if len(hs) == 1 else sum(hs)
        h = K.linear(self.concat, torch.cat([states[idx] for idx in self._concat_node], dim=1))
        return K.bn_act(h, self.batchnorm_h, relu=True)


class Network(nn.Module):
    """reference: model_lp.py:77-150"""

    def __init__(self, device, genotype, number_of_nodes, num_rels, feature_dim, init_fea_dim, num_base_r, criterion,
                 dropout_cell, args):
        super().__init__()
        self._device = device
        self._num_ent = number_of_nodes
        self._num_rel = num_rels * 2 + 1
        self._feature_dim = feature_dim
        self.num_base_r = num_base_r
        self.init_fea_dim = init_fea_dim
        self.criterion = criterion
        self.embedding_h = nn.Embedding(self._num_ent, self.init_fea_dim)
        self.embedding_e = nn.Embedding(self.num_base_r, self._feature_dim)
        self.linear_e = nn.Linear(self.init_fea_dim, self._feature_dim)
        self.rel_wt = self.get_param([self._num_rel, self.num_base_r])
        self.cells = nn.ModuleList([Cell(args, genotype[i]) for i in range(len(genotype))])
        self.score_func = MIXED_OPS_sf[genotype[-1].score_func]({'gamma': args.gamma,
                                                                 'embed_dim': args.embed_dim,
                                                                 'conve_hid_drop': args.conve_hid_drop,
                                                                 'feat_drop': args.feat_drop,
                                                                 'num_filt': args.num_filt,
                                                                 'ker_sz': args.ker_sz,
                                                                 'k_w': args.k_w,
                                                                 'k_h': args.k_h})
        self.w_rel = self.get_param([self._feature_dim, self._feature_dim])
        self._dropout = dropout_cell

    def get_param(self, shape):
        param = nn.Parameter(torch.Tensor(*shape))
        nn.init.xavier_normal_(param, gain=nn.init.calculate_gain('relu'))
        return param

    def _embed(self, g):
        """model_lp.py:124-133: entity/relation tables through every cell."""
        if getattr(g, 'part', None) is not None:
            ent_local, rel_embed = self._embed_partitioned(g)
            return D_.AllGatherRows.apply(ent_local, g.part), rel_embed
        g.require_tables(self._num_ent, self._num_rel)
        all_ent_emb = K.linear(self.linear_e, self.embedding_h.weight)  # == embedding_h(arange(N))
        rel_embed = K.matmul(self.rel_wt, self.embedding_e.weight)
        for cell in self.cells:
            all_ent_emb = cell.forward_fused(g, all_ent_emb, rel_embed)
            all_ent_emb = F.dropout(all_ent_emb, self._dropout, training=self.training)
            rel_embed = K.matmul(rel_embed, self.w_rel)
        return all_ent_emb, rel_embed

    def _embed_partitioned(self, g):
        """Destination-partitioned form of _embed: `g` holds this rank's destinations; every cell
        reads the FULL entity table (replicated for the first cell, all-gathered over NVLink for the following
        ones), writes its own rows, and normalises with statistics summed over the ranks.  Returns the LOCAL rows
        [hi-lo, D] of the final entity table and the (replicated) relation table."""
        part = g.part
        g.require_tables(self._num_ent, self._num_rel)
        table = K.linear(self.linear_e, self.embedding_h.weight)
        rel_embed = K.matmul(self.rel_wt, self.embedding_e.weight)
        with D_.use(part):
            for k, cell in enumerate(self.cells):
                if k > 0:
                    table = D_.AllGatherRows.apply(local, part)
                local = cell.forward_fused(g, table, rel_embed)
                local = F.dropout(local, self._dropout, training=self.training)
                rel_embed = K.matmul(rel_embed, self.w_rel)
        return local, rel_embed

    def _forward_lp(self, g, subj, rel):
        all_ent_emb, rel_embed = self._embed(g)
        return self.score_func(all_ent_emb, all_ent_emb[subj], rel_embed[rel])

    def forward(self, g, subj, rel):
        return self._forward_lp(g, subj, rel)

    def _loss(self, g, subj, rel, label):
        """model_lp.py:148-150.  With the DistMult scorer and BCELoss (the README
        configuration) probabilities are never materialised: fused sigmoid+BCE kernel."""
        part = getattr(g, 'part', None)
        fused_ok = isinstance(self.criterion, nn.BCELoss) and hasattr(self.score_func, 'loss') \
            and self.criterion.reduction == 'mean' and self.criterion.weight is None
        if part is not None and fused_ok:
            # entity dimension of the 1-N scores sharded like the destinations: the [B, D] query block is the only
            # tensor exchanged; `label` holds this rank's columns [lo, hi) of the [B, N] label matrix
            ent_local, rel_embed = self._embed_partitioned(g)
            sub = D_.ShardedRowSelect.apply(ent_local, subj, part)
            loss_local = self.score_func.loss(ent_local, sub, rel_embed[rel], label)
            return D_.AllReduceSum.apply(loss_local * (part.n_local / part.n_global), part)
        if part is not None and part.world > 1:
            # the replicated-loss form would hand every rank the COMPLETE table gradient, which the partitioned
            # step's gradient sum then multiplies by the world size
            raise RuntimeError("destination-partitioned training supports sf_DisMult with mean nn.BCELoss (the "
                               "entity-sharded fused loss); got "
                               f"{type(self.score_func).__name__} / {type(self.criterion).__name__}")
        if fused_ok:
            all_ent_emb, rel_embed = self._embed(g)
            return self.score_func.loss(all_ent_emb, all_ent_emb[subj], rel_embed[rel], label)
        return self.criterion(self.forward(g, subj, rel), label)
