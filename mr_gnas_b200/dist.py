"""Multi-GPU plumbing for the message-passing path (one process per GPU, torch.distributed / NCCL).

The reference is single-device (no torch.distributed call site anywhere).  Two sharding modes:

* ``allreduce_grads`` -- data parallel over 1-N query batches (LP training): every rank runs the full-graph
  message passing on its own batch of (s, r) queries; parameter gradients are averaged with ONE flat
  all-reduce per step.  Per-GPU work is fixed => weak scaling.  This is what ``bench.py --gpus N`` runs.
* ``partition_by_dst`` / ``allreduce_stats`` -- helpers for the destination-partitioned layer:
  contiguous destination ranges balanced by in-edge count (every reduction of the path is keyed by
  destination, so aggregation stays local), node embeddings all-gathered before the edge gather, BatchNorm
  column statistics summed across ranks so the result equals the single-GPU one.
"""
import torch
import torch.distributed as dist


def _staged(t, group):
    """True when the process group cannot move device memory itself (gloo: used by the CPU tests and by the
    2-rank tests on a one-GPU box, where NCCL refuses two ranks on one device): bounce through the host."""
    return t.is_cuda and dist.get_backend(group) == "gloo"


def all_reduce_(t, group=None):
    """In-place sum all-reduce of `t` over `group` (NCCL over NVLink on the GPU path)."""
    if _staged(t, group):
        h = t.detach().cpu()
        dist.all_reduce(h, group=group)
        t.copy_(h)
    else:
        dist.all_reduce(t, group=group)
    return t


def all_gather_into_(out, inp, group=None):
    if _staged(inp, group):
        ho, hi = out.detach().cpu(), inp.detach().cpu()
        dist.all_gather_into_tensor(ho, hi, group=group)
        out.copy_(ho)
    else:
        dist.all_gather_into_tensor(out, inp, group=group)
    return out


def allreduce_grads(params, world_size=None, group=None):
    """Average .grad of `params` across ranks with one flat all-reduce (deterministic packing order)."""
    world_size = world_size or (dist.get_world_size(group) if dist.is_initialized() else 1)
    if world_size <= 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    all_reduce_(flat, group)
    flat /= world_size
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def partition_by_dst(csr_ptr, num_parts):
    """Contiguous destination ranges [lo, hi) with ~equal numbers of in-edges (csr_ptr: [N+1] tensor).
    Returns a list of (node_lo, node_hi, edge_lo, edge_hi) -- edge bounds are positions in the dst-CSR."""
    ptr = csr_ptr.detach().cpu().long()
    n, total = ptr.numel() - 1, int(ptr[-1])
    bounds = [0]
    for k in range(1, num_parts):
        target = (total * k + num_parts - 1) // num_parts
        cut = int(torch.searchsorted(ptr, torch.tensor(target)).item())
        bounds.append(min(max(cut, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[k], bounds[k + 1], int(ptr[bounds[k]]), int(ptr[bounds[k + 1]])) for k in range(num_parts)]


def allreduce_stats(stats, group=None):
    """Sum per-column (sum, sum-of-squares) statistics across ranks (SyncBN for edge-row BatchNorm): the folded
    [2, D] double vector of each rank is all-reduced; mrg_bn_finalize then runs with nparts=1 and the GLOBAL
    row count so every rank derives identical scale/shift."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        all_reduce_(stats, group)
    return stats


# ------------------------------------------------------------------------------------------
# Destination-partitioned message passing
# ------------------------------------------------------------------------------------------
class Partition:
    """This rank's share of a destination-partitioned graph: destinations [lo, hi) of n_global nodes and the
    e_local of e_global directed edges that point at them.  `rows_global` maps a LOCAL BatchNorm row count to
    the global one (edge-expanded rows M = E + N, node rows N) so every rank normalises with global statistics."""

    def __init__(self, rank, world, lo, hi, n_global, e_local, e_global, bounds, group=None):
        self.rank, self.world, self.lo, self.hi = int(rank), int(world), int(lo), int(hi)
        self.n_global, self.e_local, self.e_global = int(n_global), int(e_local), int(e_global)
        self.bounds = list(bounds)          # [(lo, hi)] of every rank
        self.group = group
        self.n_local = self.hi - self.lo

    def rows_global(self, rows_local):
        """Global row count behind a local BatchNorm: edge-expanded rows (LP: E + N), edge rows (NC blocks: E) or
        node rows (N)."""
        table = {}
        for loc, glob in ((self.e_local + self.n_local, self.e_global + self.n_global),
                          (self.e_local, self.e_global), (self.n_local, self.n_global)):
            if table.setdefault(loc, glob) != glob:
                raise RuntimeError(f"ambiguous partitioned BatchNorm row count {loc} (edges {self.e_local}, nodes "
                                   f"{self.n_local}): choose another world size")
        if rows_local not in table:
            raise RuntimeError(f"partitioned BatchNorm over {rows_local} rows: not the local edge-expanded rows "
                               f"({self.e_local + self.n_local}), edges ({self.e_local}) or nodes ({self.n_local})")
        return table[rows_local]


_current = None


def current():
    """The Partition whose collectives the running forward should use (None: single-GPU semantics)."""
    return _current


class use:
    """with dist.use(g.part): ...   -- BatchNorm statistics inside are summed over the partition's ranks."""

    def __init__(self, part):
        self.part = part

    def __enter__(self):
        global _current
        self.prev, _current = _current, self.part
        return self.part

    def __exit__(self, *exc):
        global _current
        _current = self.prev
        return False


def sync_stats(part, stats, nparts, width, rows):
    """Per-CTA partial column sums ([nparts, width] doubles) of THIS rank -> ([1, width] global sums, 1,
    global row count).  The local fold is a fixed-order sum; the cross-rank sum is one all-reduce."""
    if part is None or part.world <= 1:
        return stats, nparts, rows
    folded = stats[:nparts * width].view(nparts, width).sum(0)
    all_reduce_(folded, part.group)
    return folded.contiguous(), 1, part.rows_global(rows)


def unshare_param_grads(part, *grads):
    """BatchNorm dgamma/dbeta come out of the GLOBAL statistics, i.e. complete and identical on every rank, while
    every other parameter gradient of the partitioned step is a per-rank partial that allreduce_grads_sum adds up.
    Scale the complete ones by 1/world so the same final sum restores them (exact for power-of-two worlds)."""
    if part is None or part.world <= 1:
        return grads
    return tuple(g / part.world for g in grads)


class AllGatherRows(torch.autograd.Function):
    """[n_local, D] owned rows -> the full [n_global, D] table on every rank (the halo exchange before the edge
    gather of the next cell / the entity table of the scorer).  Backward: every rank holds a partial gradient of
    the whole table (from its own edges / score block); the owner's slice of their SUM comes back."""

    @staticmethod
    def forward(ctx, x, part):
        ctx.part = part
        if part.world == 1:
            return x.clone()
        sizes = [h - l for l, h in part.bounds]
        pad = max(sizes)
        buf = x.new_zeros(pad, x.shape[1])
        buf[:x.shape[0]] = x
        out = x.new_empty(part.world * pad, x.shape[1])
        all_gather_into_(out, buf, part.group)
        return torch.cat([out[r * pad:r * pad + sizes[r]] for r in range(part.world)], 0)

    @staticmethod
    def backward(ctx, g):
        part = ctx.part
        if part.world == 1:
            return g, None
        g = g.clone(memory_format=torch.contiguous_format)   # autograd may share `g` with another consumer
        all_reduce_(g, part.group)
        return g[part.lo:part.hi].contiguous(), None


class ShardedRowSelect(torch.autograd.Function):
    """table_local[idx - lo] for the idx this rank owns, summed over ranks -> the selected rows of the global
    table on every rank ([B, D], B small: the 1-N query block of sf_DisMult, model_lp.py:135)."""

    @staticmethod
    def forward(ctx, table, idx, part):
        own = (idx >= part.lo) & (idx < part.hi)
        loc = (idx - part.lo).clamp(0, max(part.n_local - 1, 0))
        out = table[loc] * own.unsqueeze(1).to(table.dtype)
        if part.world > 1:
            all_reduce_(out, part.group)
        ctx.part, ctx.rows = part, table.shape[0]
        ctx.save_for_backward(own, loc)
        return out

    @staticmethod
    def backward(ctx, g):
        own, loc = ctx.saved_tensors
        if ctx.part.world > 1:
            g = g.clone(memory_format=torch.contiguous_format)   # never all-reduce autograd's own buffer in place
            all_reduce_(g, ctx.part.group)
        d = g.new_zeros(ctx.rows, g.shape[1])
        # static shapes (CUDA-graph capturable): rows of other ranks add exact zeros to a clamped index;
        # index_put_(accumulate) is sort-based on CUDA, hence deterministic
        d.index_put_((loc,), g * own.unsqueeze(1).to(g.dtype), accumulate=True)
        return d, None, None


class AllReduceSum(torch.autograd.Function):
    """Sum of per-rank partial losses; every rank's partial receives the upstream gradient unchanged."""

    @staticmethod
    def forward(ctx, x, part):
        y = x.clone()
        if part.world > 1:
            all_reduce_(y, part.group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def allreduce_grads_sum(params, part):
    """Partial parameter gradients (each rank saw only its own destinations / score block) -> their sum."""
    if part is None or part.world <= 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    all_reduce_(flat, part.group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def lp_partition(triples, num_ent, num_rels, rank, world, device="cuda", group=None):
    """Destination-partitioned build_graph (train/mr_lp_train.py:77-89): edges [s->o | o->s], e_type [r | r+R];
    this rank keeps the edges whose destination falls in its range (ranges balanced by in-edge count) and the
    GLOBAL degree norms.  Returns an MRGraph with .part set."""
    import numpy as np
    from .graph import MRGraph
    t = np.asarray(triples)
    s, r, o = t[:, 0].astype(np.int64), t[:, 1].astype(np.int64), t[:, 2].astype(np.int64)
    src, dst, et = np.concatenate([s, o]), np.concatenate([o, s]), np.concatenate([r, r + num_rels])
    E, T = src.shape[0], s.shape[0]
    deg = np.bincount(dst, minlength=num_ent)
    ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(deg)]))
    ranges = partition_by_dst(ptr, world)
    lo, hi = ranges[rank][0], ranges[rank][1]
    if hi <= lo:
        raise RuntimeError(f"rank {rank} of {world} would own no destination (a hub holds more than 1/{world} of the "
                           "edges): use fewer ranks for this graph")
    keep = (dst >= lo) & (dst < hi)
    half = int(keep[:T].sum())
    degf = deg.astype(np.float32)
    with np.errstate(divide="ignore"):
        n_norm = degf ** -0.5
    n_norm[np.isinf(n_norm)] = 0
    g = MRGraph.from_partition(torch.from_numpy(src[keep]), torch.from_numpy(dst[keep]), torch.from_numpy(et[keep]),
                               num_ent, 2 * num_rels + 1, lo, hi, half, torch.from_numpy(n_norm), device)
    g.part = Partition(rank, world, lo, hi, num_ent, int(keep.sum()), E, [(a, b) for a, b, _, _ in ranges], group)
    return g


def nc_partition(src, dst, etype, num_nodes, layers, rank, world, device="cuda", group=None):
    """Destination-partitioned FULL-GRAPH message-flow blocks for the NC network (models/model.py:152-184 run on the
    whole graph instead of sampled 64-seed blocks -- BASELINE configs[3]): this rank's block holds the in-edges
    of destinations [lo, hi) (ranges balanced by in-edge count), in parent-edge-id order, with parent edge ids,
    edge types and global destination ids exactly as a DGL block carries them.  Every layer uses the same block.
    Returns ([block] * layers, Partition)."""
    import numpy as np
    from .graph import MRBlock
    src, dst, etype = (np.asarray(a).astype(np.int64) for a in (src, dst, etype))
    deg = np.bincount(dst, minlength=num_nodes)
    ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(deg)]))
    ranges = partition_by_dst(ptr, world)
    lo, hi = ranges[rank][0], ranges[rank][1]
    if hi <= lo:
        raise RuntimeError(f"rank {rank} of {world} would own no destination: use fewer ranks for this graph")
    eids = np.nonzero((dst >= lo) & (dst < hi))[0]
    block = MRBlock.build(torch.from_numpy(eids), torch.from_numpy(etype[eids]), torch.from_numpy(dst[eids] - lo),
                          torch.arange(lo, hi), device)
    part = Partition(rank, world, lo, hi, num_nodes, eids.shape[0], src.shape[0], [(a, b) for a, b, _, _ in ranges],
                     group)
    block.part = part
    return [block] * layers, part
