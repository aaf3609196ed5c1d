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
