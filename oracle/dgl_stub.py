"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Pure-torch stand-in for the pieces of DGL (pinned by the reference at
``dgl-cuda10.0==0.5.3``, README.md:16) that the MR-GNAS hot path touches, so that
the reference's own Python under ``/root/reference`` can be imported *unmodified* in
the build container (no DGL wheel, no network) and used to generate the golden
vectors in ``tests/golden/`` (see ``oracle/make_golden.py``).

DGL itself is NOT vendored in the reference, so its ``update_all`` semantics are
restated here from the published behaviour of DGL 0.5.x gspmm (assumptions A1-A3,
SURVEY.md section 8 row a-8):

  A1  sum / max over a node with zero in-edges yields 0 (max: -inf replaced by 0);
  A2  mean divides the sum by max(in_degree, 1);
  A3  max records, per (dst, feature), the FIRST maximal edge in dst-CSR order, i.e.
      the lowest edge id (stable COO->CSR), and routes the gradient to that edge only.

Python-UDF reducers (reference models/operations.py:105-176) are served through a
degree-bucketing scheduler: one call per distinct in-degree with a mailbox
``[n_bucket, deg, D]`` ordered by ascending edge id; zero-in-degree nodes get zeros.

"parity unpinned": the reference ships no tests / golden vectors for this boundary,
so this stub is the only statement of DGL's behaviour the build has.
"""
import sys
import types

import torch


class _Desc:
    def __init__(self, kind, a, b):
        self.kind, self.a, self.b = kind, a, b


class _SegMax(torch.autograd.Function):
    """A3: per-(dst, feature) max with lowest-edge-id arg and single-edge gradient."""

    @staticmethod
    def forward(ctx, m, dst, num_nodes):
        E, D = m.shape
        out = torch.full((num_nodes, D), float("-inf"), dtype=m.dtype)
        arg = torch.full((num_nodes, D), -1, dtype=torch.long)
        if E > 0:
            idx = dst.view(-1, 1).expand(E, D)
            out = out.scatter_reduce(0, idx, m, reduce="amax", include_self=True)
            # lowest edge id attaining the max
            is_max = m == out[dst]
            eid = torch.arange(E).view(-1, 1).expand(E, D)
            cand = torch.where(is_max, eid, torch.full_like(eid, E))
            arg = torch.full((num_nodes, D), E, dtype=torch.long).scatter_reduce(
                0, idx, cand, reduce="amin", include_self=True)
            arg = torch.where(arg == E, torch.full_like(arg, -1), arg)
        out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
        ctx.save_for_backward(arg)
        ctx.E = E
        return out, arg

    @staticmethod
    def backward(ctx, g, _garg):
        (arg,) = ctx.saved_tensors
        E = ctx.E
        N, D = arg.shape
        dm = torch.zeros(E, D, dtype=g.dtype)
        mask = arg >= 0
        cols = torch.arange(D).view(1, -1).expand(N, D)
        dm.index_put_((arg[mask], cols[mask]), g[mask], accumulate=True)
        return dm, None, None


class _Nodes:
    def __init__(self, mailbox):
        self.mailbox = mailbox


class StubGraph:
    """Homogeneous graph / block with the attributes the reference ops and nets touch."""

    def __init__(self, num_nodes=0, src=None, dst=None):
        self._n = int(num_nodes)
        self._src = torch.zeros(0, dtype=torch.long) if src is None else torch.as_tensor(src).long()
        self._dst = torch.zeros(0, dtype=torch.long) if dst is None else torch.as_tensor(dst).long()
        self.edata, self.ndata = {}, {}
        self.dstdata = self.ndata
        self.srcdata = self.ndata
        self.last_arg = None
        self.arg_trace = []        # every update_all(max) appends its argmax table (test infrastructure)

    # --- construction API used by the scripts (mr_lp_train.py:78-88) ---
    def add_nodes(self, n):
        self._n += int(n)

    def add_edges(self, u, v):
        self._src = torch.cat([self._src, torch.as_tensor(u).long().view(-1)])
        self._dst = torch.cat([self._dst, torch.as_tensor(v).long().view(-1)])

    def number_of_nodes(self):
        return self._n

    def num_nodes(self):
        return self._n

    def num_edges(self):
        return int(self._src.numel())

    number_of_edges = num_edges

    def nodes(self):
        return torch.arange(self._n)

    def edges(self, form="uv"):
        eid = torch.arange(self.num_edges())
        if form == "all":
            return self._src, self._dst, eid
        return self._src, self._dst

    all_edges = edges

    def in_degrees(self, v=None):
        deg = torch.bincount(self._dst, minlength=self._n)
        if v is None:
            return deg
        return deg[torch.as_tensor(list(v)).long()]

    def apply_edges(self, fn):
        class _E:
            pass
        e = _E()
        e.src = {k: v[self._src] for k, v in self.ndata.items()}
        e.dst = {k: v[self._dst] for k, v in self.ndata.items()}
        e.data = self.edata
        self.edata.update(fn(e))

    def to(self, device):
        return self

    def local_var(self):
        return self

    # --- message passing ---
    def update_all(self, msg, red):
        m = self.edata[msg.a]
        N, dst = self._n, self._dst
        if callable(red) and not isinstance(red, _Desc):
            self.ndata["h"] = self._bucketed(m, red)
            return
        out_key = red.b
        if red.kind == "sum":
            out = torch.zeros(N, m.shape[1], dtype=m.dtype).index_add(0, dst, m)
        elif red.kind == "mean":
            out = torch.zeros(N, m.shape[1], dtype=m.dtype).index_add(0, dst, m)
            deg = torch.bincount(dst, minlength=N).clamp(min=1).to(m.dtype).view(-1, 1)
            out = out / deg
        elif red.kind == "max":
            out, arg = _SegMax.apply(m, dst, N)
            self.last_arg = arg
            self.arg_trace.append(arg)
        else:
            raise NotImplementedError(red.kind)
        self.ndata[out_key] = out

    def _bucketed(self, m, red):
        N, dst = self._n, self._dst
        D = m.shape[1]
        deg = torch.bincount(dst, minlength=N)
        order = torch.sort(dst, stable=True).indices  # ascending dst, ascending eid inside
        ptr = torch.zeros(N + 1, dtype=torch.long)
        ptr[1:] = torch.cumsum(deg, 0)
        out = torch.zeros(N, D, dtype=m.dtype)
        pieces, rows = [], []
        for d in torch.unique(deg).tolist():
            if d == 0:
                continue
            nodes = torch.nonzero(deg == d).view(-1)
            idx = ptr[nodes].view(-1, 1) + torch.arange(d).view(1, -1)
            mailbox = m[order[idx]]  # [n_bucket, d, D]
            pieces.append(red(_Nodes({"m": mailbox}))["h"])
            rows.append(nodes)
        if pieces:
            out = out.index_put((torch.cat(rows),), torch.cat(pieces))
        return out


def install():
    """Put stub modules into sys.modules so /root/reference imports unmodified."""
    dgl = types.ModuleType("dgl")
    dgl.__path__ = []
    dgl.DGLGraph = StubGraph
    dgl.graph = lambda data=None, **kw: StubGraph(0)
    dgl.EID, dgl.NID, dgl.ETYPE = "_ID", "_ID", "_TYPE"
    fn = types.ModuleType("dgl.function")
    fn.copy_edge = lambda e, m: _Desc("copy", e, m)
    fn.copy_e = fn.copy_edge
    fn.max = lambda m, h: _Desc("max", m, h)
    fn.sum = lambda m, h: _Desc("sum", m, h)
    fn.mean = lambda m, h: _Desc("mean", m, h)
    fn.u_sub_e = lambda u, e, out=None: _Desc("u_sub_e", u, e)
    fn.u_mul_e = lambda u, e, out=None: _Desc("u_mul_e", u, e)
    dgl.function = fn
    data = types.ModuleType("dgl.data")
    data.__path__ = []
    rdf = types.ModuleType("dgl.data.rdf")
    for n in ("AIFBDataset", "MUTAGDataset", "BGSDataset", "AMDataset"):
        setattr(rdf, n, object)
    data.rdf = rdf
    dgl.data = data
    sys.modules.update({"dgl": dgl, "dgl.function": fn, "dgl.data": data, "dgl.data.rdf": rdf})
    gml = types.ModuleType("utils.gpu_memory_log")
    gml.gpu_memory_log = lambda *a, **k: None
    sys.modules["utils.gpu_memory_log"] = gml
    return dgl
