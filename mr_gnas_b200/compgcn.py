"""CompGCN layer (reference: models/compgcn.py:12-113, an orphan mirror of DGL's example) on libmrgnas.
The composition runs in the K1 kernel, the per-direction transforms are edge-tile GEMMs on the masked rows,
``update_all(copy_e, sum)`` is the deterministic segmented-sum kernel.  ``ccorr`` (circular correlation)
uses torch.fft because the reference's torch.rfft (utils/utils.py:301) no longer exists."""
import torch
import torch.nn as nn

from . import functional as K


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
        src, _ = g.edges()
        e_h = (r_feats[etype.long()] * g.edata['norm'].view(-1, 1)).contiguous()
        comp_h = self._comp(n_in_feats[src].contiguous(), e_h)
        in_idx = torch.nonzero(g.edata['in_edges_mask'], as_tuple=False).view(-1)
        out_idx = torch.nonzero(g.edata['out_edges_mask'], as_tuple=False).view(-1)
        new_comp_h = torch.zeros(comp_h.shape[0], self.out_dim, device=comp_h.device)
        new_comp_h = new_comp_h.index_put((out_idx,), self.W_O(comp_h[out_idx]))
        new_comp_h = new_comp_h.index_put((in_idx,), self.W_I(comp_h[in_idx]))
        comp_edge = K.SegReduce.apply(new_comp_h, None, g, 0, False)   # update_all(copy_e, sum)
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
