"""Destination-partitioned message passing against the single-GPU path on the same seeded
inputs: loss, every parameter gradient and the BatchNorm running statistics must agree.  world=1 exercises
mrg_graph_build_part, the sharded scorer and the statistics plumbing; world=2 runs one rank per GPU over NCCL when
the box has 2 GPUs and otherwise BOTH ranks on the one GPU with a gloo group (dist.py stages the collectives
through the host there) -- the partition arithmetic is the same, so the 2-rank path is never skipped."""
import os
import types
from collections import namedtuple

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

pytestmark = pytest.mark.gpu
Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func")
CELL = Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_sparse_comp', 2, 1), ('f_sparse_comp', 3, 2), ('a_max', 4, 2),
                            ('a_max', 5, 3), ('f_sparse_last', 6, 5), ('f_sparse_last', 7, 5)],
                concat_node=[4, 5, 6, 7], score_func='sf_DisMult')


def _free_port():
    """A rendezvous FILE (not a TCP port: a freshly released port can still be in TIME_WAIT -> EADDRINUSE)."""
    import tempfile
    fd, path = tempfile.mkstemp(prefix="mrg_rdzv_")
    os.close(fd)
    os.unlink(path)
    return path


def _args(D):
    return types.SimpleNamespace(feature_dim=D, drop_aggr=0.0, drop_op=0.0, gamma=40, embed_dim=D,
                                 conve_hid_drop=0.0, feat_drop=0.0, num_filt=4, ker_sz=3, k_w=4, k_h=D // 4)


from parity import check_grads_pair, rel_err as _err  # max|a-b| / max|b|: relative to the tensor's own scale, no absolute floor


def _compare(rank, world, n_cells, dev_index=None):
    from mr_gnas_b200 import dist as D_
    from mr_gnas_b200.graph import MRGraph
    from mr_gnas_b200.model_lp import Network
    from mr_gnas_b200.synth import synth_kg
    from mr_gnas_b200.utils import weights_init
    dev = torch.device("cuda", rank if dev_index is None else dev_index)
    N, R, T, D, B = 700, 9, 6000, 64, 32
    trip = synth_kg(N, R, T, seed=3)
    genos = [CELL] * n_cells

    def fresh():
        torch.manual_seed(0)
        m = Network(dev, genos, N, R, D, D, 2 * R + 1, nn.BCELoss(), 0.0, _args(D))
        m.apply(weights_init)
        return m.to(dev).train()

    rng = np.random.RandomState(5)
    subj = torch.from_numpy(rng.randint(0, N, B)).to(dev)
    rel = torch.from_numpy(rng.randint(0, 2 * R, B)).to(dev)
    label = (torch.from_numpy(rng.rand(B, N)) < 0.02).float().to(dev) * 0.9 + 1.0 / N

    ref = fresh()
    g_full = MRGraph.from_triples(N, trip, R, device=dev)
    loss_ref = ref._loss(g_full, subj, rel, label)
    loss_ref.backward()

    par = fresh()
    g = D_.lp_partition(trip, N, R, rank, world, device=dev)
    assert g.part.lo < g.part.hi and g.E == g.part.e_local
    loss = par._loss(g, subj, rel, label[:, g.part.lo:g.part.hi].contiguous())
    loss.backward()
    D_.allreduce_grads_sum(list(par.parameters()), g.part)
    assert _err(loss, loss_ref) <= 1e-5, (float(loss), float(loss_ref))
    check_grads_pair(f"LP partition world {world}", {k: p.grad for k, p in par.named_parameters()},
                     {k: q.grad for k, q in ref.named_parameters()})
    for (k, a), b in zip(par.named_buffers(), ref.buffers()):
        if a.dtype.is_floating_point:
            assert _err(a, b) <= 1e-5, k
    # the all-gathered forward (predict path) equals the single-GPU probabilities
    par.eval(), ref.eval()
    with torch.no_grad():
        assert _err(par(g, subj, rel), ref(g_full, subj, rel)) <= 1e-5


def _init(rank, world, port):
    """-> device index of this rank.  One GPU per rank + NCCL when possible, else a shared GPU + gloo."""
    init = f"file://{port}"
    if torch.cuda.device_count() >= world:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=init, rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
        return rank
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method=init, rank=rank, world_size=world)
    return 0


def _worker(rank, world, port, n_cells):
    dev_index = _init(rank, world, port)
    try:
        _compare(rank, world, n_cells, dev_index)
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("n_cells", [1, 2])
def test_partitioned_lp_world1_matches_full_graph(n_cells):
    mp.spawn(_worker, args=(1, _free_port(), n_cells), nprocs=1, join=True)


@pytest.mark.parametrize("n_cells", [1, 2])
def test_partitioned_lp_world2_matches_single_gpu(n_cells):
    mp.spawn(_worker, args=(2, _free_port(), n_cells), nprocs=2, join=True)


# ------------------------------------------------------------------------------ NC full-graph layers, partitioned
NC_GENO = ("[Genotype(alpha_cell=[('pre_sub', 1, 0), ('f_dense', 2, 1), ('f_sparse', 3, 2), ('f_identity', 4, 3), "
           "('a_sum', 5, 2), ('a_sum', 6, 3), ('a_mean', 7, 4), ('f_dense_last', 8, 7), ('f_sparse_last', 9, 7), "
           "('f_sparse_last', 10, 5)], concat_node=[5, 6, 7, 8, 9, 10]), Genotype(alpha_cell=[('pre_sub', 1, 0), "
           "('f_sparse', 2, 1), ('f_identity', 3, 2), ('f_identity', 4, 1), ('a_max', 5, 2), ('a_mean', 6, 3), "
           "('a_mean', 7, 4), ('f_sparse_last', 8, 7), ('f_sparse_last', 9, 8), ('f_identity', 10, 9)], "
           "concat_node=[5, 6, 7, 8, 9, 10])]")


def _compare_nc(rank, world, dev_index=None):
    from mr_gnas_b200 import dist as D_
    from mr_gnas_b200.graph import MRBlock
    from mr_gnas_b200.model import Network
    GenoNC = namedtuple("Genotype", "alpha_cell concat_node score_func", defaults=(None,))
    dev = torch.device("cuda", rank if dev_index is None else dev_index)
    N, ET, E, D, D0, C, NB = 900, 8, 9000, 32, 16, 4, 5
    rng = np.random.RandomState(11)
    src, dst, et = rng.randint(0, N, E), (rng.zipf(1.6, E) - 1) % N, rng.randint(0, ET, E)
    dst[:50] = 3                                     # a hub; some nodes stay isolated
    trip_index = torch.from_numpy(np.stack([np.arange(E), src, dst], 1)).to(dev)
    labels = torch.from_numpy(rng.randint(0, C, N)).to(dev)
    idx = torch.from_numpy(rng.choice(N, 120, replace=False)).to(dev)
    args = types.SimpleNamespace(feature_dim=D, op_norm=True)
    genos = eval(NC_GENO, {"Genotype": GenoNC})

    def fresh():
        torch.manual_seed(0)
        return Network(dev, genos, N, C, ET, 2, 1, 2, D, D0, NB, nn.CrossEntropyLoss(), args).to(dev).train()

    ref = fresh()
    full = MRBlock.build(torch.arange(E), torch.from_numpy(et), torch.from_numpy(dst), torch.arange(N), dev)
    out_ref = ref._forward(trip_index, [full, full])
    loss_ref = nn.functional.cross_entropy(ref.classifier(out_ref)[idx], labels[idx])
    loss_ref.backward()

    par = fresh()
    blocks, part = D_.nc_partition(src, dst, et, N, 2, rank, world, dev)
    out = par._forward(trip_index, blocks)
    assert _err(out, out_ref[part.lo:part.hi]) <= 1e-5
    par.zero_grad()
    loss = par._loss_partitioned(trip_index, blocks, labels, idx)
    loss.backward()
    D_.allreduce_grads_sum(list(par.parameters()), part)
    assert _err(loss, loss_ref) <= 1e-5, (float(loss), float(loss_ref))
    check_grads_pair(f"NC partition world {world}", {k: p.grad for k, p in par.named_parameters()},
                     {k: q.grad for k, q in ref.named_parameters() if dict(par.named_parameters())[k].grad is not None})


def _worker_nc(rank, world, port):
    dev_index = _init(rank, world, port)
    try:
        _compare_nc(rank, world, dev_index)
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_partitioned_nc_full_graph_matches_single_gpu(world):
    mp.spawn(_worker_nc, args=(world, _free_port()), nprocs=world, join=True)
