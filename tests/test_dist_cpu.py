"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (flat gradient all-reduce used by
bench.py --gpus N, destination partitioning, statistics all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mr_gnas_b200.dist import allreduce_grads, allreduce_stats
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    torch.manual_seed(100 + rank)
    x, y = torch.randn(7, 6), torch.randn(7, 3)
    ((model(x) - y) ** 2).mean().backward()
    params = list(model.parameters())
    local = [p.grad.clone() for p in params]
    allreduce_grads(params, world)
    stats = torch.full((2, 4), float(rank + 1), dtype=torch.float64)
    allreduce_stats(stats)
    torch.save({"local": local, "avg": [p.grad.clone() for p in params], "stats": stats}, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_flat_grad_allreduce_gloo_world2(tmp_path):
    world, out = 2, str(tmp_path / "r")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = [torch.load(f"{out}.{r}") for r in range(world)]
    for k in range(len(res[0]["local"])):
        expect = (res[0]["local"][k] + res[1]["local"][k]) / 2
        assert torch.allclose(res[0]["avg"][k], expect, atol=1e-7)
        assert torch.equal(res[0]["avg"][k], res[1]["avg"][k])   # every rank ends with identical gradients
    assert torch.equal(res[0]["stats"], torch.full((2, 4), 3.0, dtype=torch.float64))


def test_partition_by_dst_is_balanced_and_covers():
    from mr_gnas_b200.dist import partition_by_dst
    torch.manual_seed(0)
    deg = torch.randint(0, 50, (1000,))
    deg[7] = 5000                       # a hub
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(deg, 0)])
    for parts in (1, 2, 4, 8):
        P = partition_by_dst(ptr, parts)
        assert P[0][0] == 0 and P[-1][1] == 1000 and P[0][2] == 0 and P[-1][3] == int(ptr[-1])
        for a, b in zip(P[:-1], P[1:]):
            assert a[1] == b[0] and a[3] == b[2]
        sizes = [p[3] - p[2] for p in P]
        assert max(sizes) <= int(ptr[-1]) / parts + 5000 + 50


def _worker_part(rank, world, port, out):
    """gloo, CPU tensors: the autograd-aware collectives of the destination-partitioned path against the
    single-process result (the CUDA kernels themselves are covered by tests/test_gpu_partition.py)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mr_gnas_b200 import dist as D_
    bounds = [(0, 5), (5, 12)]
    lo, hi = bounds[rank]
    part = D_.Partition(rank, world, lo, hi, 12, 30 + rank, 61, bounds)
    torch.manual_seed(1)
    table = torch.randn(12, 4, dtype=torch.float64)
    w = torch.randn(12, 4, dtype=torch.float64)
    idx = torch.tensor([0, 7, 7, 11, 4, 5])
    q = torch.randn(6, 4, dtype=torch.float64)
    # single-process reference: loss = sum(w * table) + sum(q * table[idx])
    t_ref = table.clone().requires_grad_(True)
    ((w * t_ref).sum() + (q * t_ref[idx]).sum()).backward()
    # partitioned: every rank owns rows [lo, hi); the gathered table feeds a per-rank PARTIAL loss
    local = table[lo:hi].clone().requires_grad_(True)
    full = D_.AllGatherRows.apply(local, part)
    sel = D_.ShardedRowSelect.apply(local, idx, part)
    mine = torch.zeros(12, dtype=torch.float64)
    mine[lo:hi] = 1
    partial = (w * full * mine.unsqueeze(1)).sum() + (q * sel).sum() / world
    total = D_.AllReduceSum.apply(partial, part)
    total.backward()
    stats = torch.arange(3 * 8, dtype=torch.float64).view(3, 8) * (rank + 1)
    folded, nparts, rows = D_.sync_stats(part, stats.reshape(-1), 3, 8, part.e_local + part.n_local)
    torch.save({"full": full.detach(), "sel": sel.detach(), "total": total.detach(), "grad": local.grad, "lo": lo,
                "hi": hi, "ref_grad": t_ref.grad, "ref_total": (w * table).sum() + (q * table[idx]).sum(),
                "folded": folded, "nparts": nparts, "rows": rows, "table": table, "idx": idx}, f"{out}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_partition_collectives_gloo_world2(tmp_path):
    world, out = 2, str(tmp_path / "p")
    mp.spawn(_worker_part, args=(world, _free_port(), out), nprocs=world, join=True)
    res = [torch.load(f"{out}.{r}") for r in range(world)]
    for r in res:
        assert torch.equal(r["full"], r["table"])                       # halo all-gather restores the table
        assert torch.allclose(r["sel"], r["table"][r["idx"]])           # sharded row select == plain indexing
        assert torch.allclose(r["total"], r["ref_total"])
        assert torch.allclose(r["grad"], r["ref_grad"][r["lo"]:r["hi"]])  # owner's slice of the summed gradient
        assert r["nparts"] == 1 and r["rows"] == 61 + 12
        expect = (torch.arange(24, dtype=torch.float64).view(3, 8).sum(0)) * 3   # ranks contribute x1 and x2
        assert torch.equal(r["folded"], expect)


def test_lp_partition_split_is_exact():
    """Host side of lp_partition: every directed edge lands on exactly one rank, in global edge-id order, with
    its direction half preserved."""
    import numpy as np
    from mr_gnas_b200.dist import partition_by_dst
    from mr_gnas_b200.synth import synth_kg
    N, R, T = 300, 5, 2000
    t = np.asarray(synth_kg(N, R, T, seed=1))
    s, o = t[:, 0], t[:, 2]
    dst = np.concatenate([o, s])
    deg = np.bincount(dst, minlength=N)
    ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(deg)]))
    for world in (1, 2, 3, 8):
        ranges = partition_by_dst(ptr, world)
        owner = np.full(2 * T, -1)
        for r, (lo, hi, e_lo, e_hi) in enumerate(ranges):
            keep = (dst >= lo) & (dst < hi)
            assert keep.sum() == e_hi - e_lo
            assert (owner[keep] == -1).all()
            owner[keep] = r
        assert (owner >= 0).all()
