"""MRGraph: the device-resident multi-relational graph that replaces the DGL graph object on
the hot path.  It owns the K0 arrays (dst-CSR, src-CSC, relation segments, chunk tables,
degree norms) built by ``mrg_graph_build`` and offers the handful of DGL methods the
reference's operators, networks and scripts touch:

  ops      : g.num_edges(), g.edata['norm'], g.dstdata / g.ndata
  networks : g.edges(form='all'), g.nodes(), g.edata['e_type']      (model_lp.py:126-129)
  scripts  : add_nodes / add_edges / in_degrees / number_of_nodes / ndata / apply_edges /
             to(device) / local_var()                                (mr_lp_train.py:77-89)

Row-order contract (reference, operations_lp.py:318-337): edge id i <-> row i; rows
[0,E/2) original direction, [E/2,E) inverse, rows [E,E+N) self loops with relation 2R.
"""
import ctypes

import numpy as np
import torch

from . import _lib


class _Segments:
    """ptr/idx segment list + its chunk table (all int32 on device)."""

    def __init__(self, ptr, idx, nseg, total, device):
        self.ptr, self.idx, self.nseg, self.total = ptr, idx, int(nseg), int(total)
        lib = _lib.load()
        self.max_chunks = int(lib.mrg_chunk_capacity(self.total, self.nseg))
        self.chunk_first = torch.empty(self.nseg + 1, dtype=torch.int32, device=device)
        self.chunk_seg = torch.empty(max(self.max_chunks, 1), dtype=torch.int32, device=device)
        wsb = int(lib.mrg_chunk_workspace_bytes(self.nseg))
        ws = torch.empty(wsb, dtype=torch.uint8, device=device)
        _lib.call("mrg_chunk_build", _lib.ptr(ptr), self.nseg, _lib.ptr(self.chunk_first), _lib.ptr(self.chunk_seg),
                  _lib.ptr(ws), wsb, _lib.stream())
        self._ws = {}

    def workspace(self, D, kind):
        key = (D, kind == 2)
        if key not in self._ws:
            nbytes = int(_lib.load().mrg_seg_reduce_workspace_bytes(self.max_chunks, D, kind))
            self._ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=self.ptr.device)
        return self._ws[key]


class MRGraph:
    def __init__(self, num_nodes=0, device=None):
        # deferred-build mode used by the reference scripts: DGLGraph(); add_nodes; add_edges...
        self._n = int(num_nodes)
        self._pending_src, self._pending_dst = [], []
        self._device = torch.device(device) if device is not None else None
        self._built = False
        self.edata, self.ndata = {}, {}
        self.dstdata = self.ndata
        self.srcdata = self.ndata
        self.last_arg = None
        self.n_rel_rows = None
        self.half = None      # rows [0, half) original direction, [half, E) inverse (E // 2 for the full graph)
        self.part = None      # dist.Partition when this graph holds one destination range of a larger graph
        self.n_src = None     # rows of the gather tables (== N unless partitioned)
        self._rel_rows_inferred = False

    # ----------------------------------------------------------------- construction
    @classmethod
    def from_edges(cls, src, dst, etype, num_nodes, n_rel_rows, device="cuda", with_norm=True):
        g = cls(num_nodes, device)
        g._finalize(torch.as_tensor(src), torch.as_tensor(dst), torch.as_tensor(etype), int(n_rel_rows), with_norm)
        return g

    @classmethod
    def from_triples(cls, num_ent, triples, num_rels, device="cuda"):
        """edges [s->o | o->s], e_type [r | r+R] -- train/mr_lp_train.py:77-89 (build_graph)."""
        t = torch.as_tensor(np.asarray(triples))
        s, r, o = t[:, 0], t[:, 1], t[:, 2]
        return cls.from_edges(torch.cat([s, o]), torch.cat([o, s]), torch.cat([r, r + num_rels]), num_ent,
                              2 * num_rels + 1, device)

    @classmethod
    def from_partition(cls, src, dst, etype, num_nodes, n_rel_rows, node_lo, node_hi, half, n_norm, device="cuda"):
        """The edges (global edge-id order, global node ids) that point into destinations [node_lo, node_hi) of a
        num_nodes-node graph; `half` of them belong to the original direction; n_norm [num_nodes] are the GLOBAL
        in-degree norms (mr_lp_train.py:81-84).  Destinations and outputs are local, gather tables global."""
        dev = torch.device(device)
        g = cls(int(node_hi - node_lo), dev)
        lib = _lib.load()
        E, n_src, n_dst = int(src.numel()), int(num_nodes), int(node_hi - node_lo)
        i32 = dict(dtype=torch.int32, device=dev)
        g.E, g.N, g.M, g.n_rel_rows, g.half, g.n_src = E, n_dst, E + n_dst, int(n_rel_rows), int(half), n_src
        g.src = torch.as_tensor(src).to(device=dev, dtype=torch.int32).contiguous()
        dst_g = torch.as_tensor(dst).to(device=dev, dtype=torch.int32).contiguous()
        g.dst = (dst_g - int(node_lo)).contiguous()
        g.etype = torch.as_tensor(etype).to(device=dev, dtype=torch.int32).contiguous()
        csr_ptr, csr_eid = torch.empty(n_dst + 1, **i32), torch.empty(max(E, 1), **i32)
        csc_ptr, csc_row = torch.empty(n_src + 1, **i32), torch.empty(g.M, **i32)
        rel_ptr, rel_row = torch.empty(g.n_rel_rows + 1, **i32), torch.empty(g.M, **i32)
        wsb = int(lib.mrg_graph_workspace_bytes(E, n_dst, g.n_rel_rows))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        p = _lib.ptr
        _lib.call("mrg_graph_build_part", p(g.src), p(g.dst), p(g.etype), E, n_src, n_dst, int(node_lo), g.n_rel_rows,
                  p(csr_ptr), p(csr_eid), p(csc_ptr), p(csc_row), p(rel_ptr), p(rel_row), p(ws), wsb, _lib.stream())
        g.in_deg = (csr_ptr[1:] - csr_ptr[:-1]).contiguous()
        g.n_norm = torch.as_tensor(n_norm).to(device=dev, dtype=torch.float32).contiguous()
        g.edge_norm = torch.empty(E, dtype=torch.float32, device=dev)
        _lib.call("mrg_edge_norm", p(g.src), p(dst_g), p(g.n_norm), E, p(g.edge_norm), _lib.stream())
        g.csr = _Segments(csr_ptr, csr_eid, n_dst, E, dev)
        g.csc = _Segments(csc_ptr, csc_row, n_src, g.M, dev)
        g.rel = _Segments(rel_ptr, rel_row, g.n_rel_rows, g.M, dev)
        g.src_final = torch.cat([g.src, torch.arange(int(node_lo), int(node_hi), **i32)])
        g.et_final = torch.cat([g.etype, torch.full((n_dst,), g.n_rel_rows - 1, **i32)])
        g.edata["norm"] = g.edge_norm
        g.edata["e_type"] = g.etype.long()
        g._built = True
        return g

    @classmethod
    def from_block(cls, dst, num_dst, device="cuda"):
        """Bipartite block (NC path, model.py:156-174): only the dst-CSR is needed."""
        g = cls(num_dst, device)
        dst = torch.as_tensor(dst)
        g._finalize(torch.zeros_like(dst), dst, torch.zeros_like(dst), 1, with_norm=False, dst_only=True)
        return g

    def add_nodes(self, n):
        self._n += int(n)

    def add_edges(self, u, v):
        self._pending_src.append(torch.as_tensor(np.asarray(u) if not torch.is_tensor(u) else u).long().view(-1))
        self._pending_dst.append(torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).long().view(-1))
        self._built = False

    def to(self, device):
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise RuntimeError("MRGraph lives on a CUDA device: the message-passing path has no CPU fallback")
        self._ensure()
        for d in (self.edata, self.ndata):
            for k, v in list(d.items()):
                if torch.is_tensor(v):
                    d[k] = v.to(self._device)
        return self

    def local_var(self):
        return self

    def _ensure(self):
        if self._built:
            return
        if self._device is None or self._device.type != "cuda":
            return  # still being assembled on the host by the script
        src = torch.cat(self._pending_src) if self._pending_src else torch.zeros(0, dtype=torch.long)
        dst = torch.cat(self._pending_dst) if self._pending_dst else torch.zeros(0, dtype=torch.long)
        et = self.edata.get("e_type")
        if et is None:
            et = torch.zeros_like(src)
            nrel = 1
        else:
            nrel = int(et.max().item()) + 2 if et.numel() else 1  # + self-loop relation row
        # The DGL-style deferred build never states the relation count: what the data shows is a lower bound
        # (the highest inverse relation may have no edge).  The model states the real table size through
        # require_tables(), which rebuilds the relation segments if the guess was short.
        self._rel_rows_inferred = True
        self._finalize(src, dst, et, nrel, with_norm=True)

    def require_tables(self, n_ent_rows, n_rel_rows):
        """Called by the networks before the gather: the entity / relation tables must have exactly the rows this
        graph's gather indices and backward segment lists cover (src ids < n_src, relation ids < n_rel_rows, the
        self-loop rows use relation n_rel_rows - 1 = 2R).  A graph assembled DGL-style (add_edges + edata) only
        guessed its relation count and is re-segmented here; anything else is an error, not silent garbage."""
        self._ensure()
        if int(n_ent_rows) != self.n_src:
            raise RuntimeError(f"entity table has {n_ent_rows} rows, the graph gathers from {self.n_src}")
        if int(n_rel_rows) == self.n_rel_rows:
            return self
        if self._rel_rows_inferred and self.part is None and int(n_rel_rows) > int(self.etype.max().item()) + 1:
            norm = self.edata.get("norm")
            self._finalize(self.src, self.dst, self.etype, int(n_rel_rows), with_norm=True)
            if norm is not None:         # keep a norm the script wrote itself (mr_lp_search.py:30-36)
                self.edata["norm"] = norm
            return self
        raise RuntimeError(f"relation table has {n_rel_rows} rows, the graph was built for {self.n_rel_rows}")

    def _finalize(self, src, dst, etype, n_rel_rows, with_norm=True, dst_only=False):
        dev = self._device
        lib = _lib.load()
        E, N = int(src.numel()), self._n
        M = E + N
        i32 = dict(dtype=torch.int32, device=dev)
        self.E, self.N, self.M, self.n_rel_rows = E, N, M, int(n_rel_rows)
        self.half, self.n_src = E // 2, N
        self.src = src.to(device=dev, dtype=torch.int32).contiguous()
        self.dst = dst.to(device=dev, dtype=torch.int32).contiguous()
        self.etype = etype.to(device=dev, dtype=torch.int32).contiguous()
        self.in_deg = torch.empty(N, **i32)
        self.n_norm = torch.empty(N, dtype=torch.float32, device=dev) if with_norm else None
        self.edge_norm = torch.empty(E, dtype=torch.float32, device=dev) if with_norm else None
        csr_ptr, csr_eid = torch.empty(N + 1, **i32), torch.empty(max(E, 1), **i32)
        if dst_only:
            csc_ptr = csc_row = rel_ptr = rel_row = None
        else:
            csc_ptr, csc_row = torch.empty(N + 1, **i32), torch.empty(M, **i32)
            rel_ptr, rel_row = torch.empty(self.n_rel_rows + 1, **i32), torch.empty(M, **i32)
        wsb = int(lib.mrg_graph_workspace_bytes(E, N, self.n_rel_rows))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        p = _lib.ptr
        _lib.call("mrg_graph_build", p(self.src), p(self.dst), p(self.etype), E, N, self.n_rel_rows, p(self.in_deg),
                  p(self.n_norm), p(self.edge_norm), p(csr_ptr), p(csr_eid), p(csc_ptr), p(csc_row), p(rel_ptr),
                  p(rel_row), p(ws), wsb, _lib.stream())
        if with_norm:
            # n_norm exactly as the reference computes it: numpy float32 in_deg ** -0.5, inf -> 0
            # (train/mr_lp_train.py:81-84); N floats, one-off, so the host does it and the result is
            # bit-identical to the reference's.  The per-edge product is one fp32 multiply on the device.
            deg = self.in_deg.cpu().numpy().astype(np.float32)
            with np.errstate(divide="ignore"):
                nn_ = deg ** -0.5
            nn_[np.isinf(nn_)] = 0
            self.n_norm = torch.from_numpy(nn_).to(dev)
            _lib.call("mrg_edge_norm", p(self.src), p(self.dst), p(self.n_norm), E, p(self.edge_norm), _lib.stream())
        self.csr = _Segments(csr_ptr, csr_eid, N, E, dev)
        if not dst_only:
            self.csc = _Segments(csc_ptr, csc_row, N, M, dev)
            self.rel = _Segments(rel_ptr, rel_row, self.n_rel_rows, M, dev)
            # edge-expanded row -> (source entity, relation row)  (model_lp.py:126-129)
            self.src_final = torch.cat([self.src, torch.arange(N, **i32)])
            self.et_final = torch.cat([self.etype, torch.full((N,), self.n_rel_rows - 1, **i32)])
        if with_norm:
            self.edata["norm"] = self.edge_norm
            self.ndata["n_norm"] = self.n_norm
        self.edata["e_type"] = self.etype.long()
        self._built = True
        del ws

    # ----------------------------------------------------------------- DGL-shaped accessors
    def number_of_nodes(self):
        return self._n

    num_nodes = number_of_nodes

    def num_edges(self):
        if self._built:
            return self.E
        return int(sum(t.numel() for t in self._pending_src))

    number_of_edges = num_edges

    def nodes(self):
        return torch.arange(self._n, device=self._device if self._built else None)

    def edges(self, form="uv"):
        self._ensure()
        s, d = self.src.long(), self.dst.long()
        if form == "all":
            return s, d, torch.arange(self.E, device=s.device)
        return s, d

    all_edges = edges

    def in_degrees(self, v=None):
        if self._built:
            deg = self.in_deg.long()
        else:
            dst = torch.cat(self._pending_dst) if self._pending_dst else torch.zeros(0, dtype=torch.long)
            deg = torch.bincount(dst, minlength=self._n)
        return deg if v is None else deg[torch.as_tensor(list(v), device=deg.device).long()]

    def apply_edges(self, fn):
        """Host-side UDF used once at graph-build time by the scripts (mr_lp_train.py:86)."""
        if self._built:
            src, dst = self.src.long(), self.dst.long()
        else:
            src, dst = torch.cat(self._pending_src), torch.cat(self._pending_dst)

        class _E:
            pass

        e = _E()
        e.src = {k: v[src] for k, v in self.ndata.items() if torch.is_tensor(v)}
        e.dst = {k: v[dst] for k, v in self.ndata.items() if torch.is_tensor(v)}
        e.data = self.edata
        self.edata.update(fn(e))

    def norm(self):
        """Per-edge scale [E] fp32 contiguous (g.edata['norm'], [E] or [E,1])."""
        n = self.edata["norm"]
        if n.dtype != torch.float32 or not n.is_contiguous() or n.dim() != 1:
            n = n.reshape(-1).float().contiguous()
        return n


class MRBlock(MRGraph):
    """One message-flow block of the NC path (DGL ``blocks[i]`` of MultiLayerFullNeighborSampler,
    train/mr_nc_train.py:42-51): E_b edges from global source nodes into n_dst local destinations.
    Exposes what models/model.py:156-178 reads: ``edata['_ID']`` (dgl.EID: parent edge ids),
    ``edata['_TYPE']`` (dgl.ETYPE) and ``dstdata['_ID']`` (dgl.NID: global ids of the destinations)."""

    @classmethod
    def build(cls, parent_eid, etype, dst_local, dst_nid, device="cuda"):
        b = cls(int(dst_nid.numel()), device)
        dst_local = torch.as_tensor(dst_local)
        b._finalize(torch.zeros_like(dst_local), dst_local, torch.zeros_like(dst_local), 1, with_norm=False,
                    dst_only=True)
        b.edata['_ID'] = torch.as_tensor(parent_eid).to(device).long()
        b.edata['_TYPE'] = torch.as_tensor(etype).to(device).long()
        b.dstdata = {'_ID': torch.as_tensor(dst_nid).to(device).long()}
        b.part = None
        return b


def full_neighbor_blocks(src, dst, etype, seeds, num_layers, device="cuda"):
    """All in-edges of the seeds, then of their sources, ... (MultiLayerFullNeighborSampler with
    return_eids=True).  Returns blocks ordered input-side first, like DGL.  Host-side glue (numpy)."""
    src, dst, etype = (np.asarray(a) for a in (src, dst, etype))
    order = np.argsort(dst, kind="stable")
    ptr = np.zeros(int(max(src.max(initial=0), dst.max(initial=0))) + 2, dtype=np.int64)
    np.add.at(ptr, dst + 1, 1)
    ptr = np.cumsum(ptr)
    blocks = []
    frontier = np.asarray(seeds, dtype=np.int64)
    for _ in range(num_layers):
        eids = np.concatenate([order[ptr[n]:ptr[n + 1]] for n in frontier]) if len(frontier) else np.zeros(0, np.int64)
        eids = np.sort(eids)  # ascending parent edge id == DGL block edge order
        local = np.searchsorted(frontier, dst[eids]) if np.all(np.diff(frontier) > 0) else \
            np.array([int(np.nonzero(frontier == d)[0][0]) for d in dst[eids]], dtype=np.int64)
        blocks.append(MRBlock.build(eids, etype[eids], local, torch.from_numpy(frontier), device))
        # next (outer) layer: destinations = seeds of this layer followed by their new sources
        nxt = np.unique(np.concatenate([frontier, src[eids]]))
        frontier = nxt
    return blocks[::-1]
