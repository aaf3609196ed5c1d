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


def allreduce_grads(params, world_size=None, group=None):
    """Average .grad of `params` across ranks with one flat all-reduce (deterministic packing order)."""
    world_size = world_size or (dist.get_world_size(group) if dist.is_initialized() else 1)
    if world_size <= 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
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
        dist.all_reduce(stats, group=group)
    return stats
