"""TEST INFRASTRUCTURE ONLY.  Compact, reproducible summaries of large tensors, so that results of the REAL
reference at BASELINE's full configuration shapes (hundreds of MB of gradients) can be committed as small
fixtures: per tensor the shape, 2-norm, max |x|, sum and the values at `k` seeded pseudo-random positions
(the whole tensor when it has <= k elements).  Positions are regenerated from the seed at test time."""
import zlib

import torch

K = 2048


def positions(numel, seed, k=K):
    if numel <= k:
        return torch.arange(numel)
    g = torch.Generator().manual_seed(int(seed))
    return torch.randint(numel, (k,), generator=g)


def seed_of(name):
    return zlib.crc32(name.encode()) & 0x7FFFFFFF


def summarize(name, t, k=K):
    flat = t.detach().cpu().reshape(-1)
    d = flat.double()
    pos = positions(flat.numel(), seed_of(name), k)
    return {"shape": tuple(t.shape), "norm": float(d.norm()), "absmax": float(d.abs().max()) if d.numel() else 0.0,
            "sum": float(d.sum()), "vals": flat[pos].clone(), "k": k}


def sample(name, t, k=K):
    flat = t.detach().cpu().reshape(-1)
    return flat[positions(flat.numel(), seed_of(name), k)]


def errors(name, t, summ):
    """-> (max |sampled diff| / absmax, |norm diff| / norm) of tensor `t` against a stored summary."""
    assert tuple(t.shape) == tuple(summ["shape"]), f"{name}: shape {tuple(t.shape)} vs {summ['shape']}"
    got = sample(name, t, summ["k"]).double()
    ref = summ["vals"].double()
    scale = summ["absmax"]
    e_val = float((got - ref).abs().max()) if got.numel() else 0.0
    e_val = e_val if scale == 0.0 else e_val / scale
    nrm = float(t.detach().double().norm())
    e_nrm = abs(nrm - summ["norm"]) if summ["norm"] == 0.0 else abs(nrm - summ["norm"]) / summ["norm"]
    return e_val, e_nrm


def checksum(a):
    """order-sensitive integer checksum of an index array (int64 arithmetic mod 2^61-1)."""
    import numpy as np
    a = np.asarray(a).astype(np.int64).reshape(-1)
    w = (np.arange(a.shape[0], dtype=np.int64) % 1000003) + 1
    return int(((a % 2147483647) * w % 2305843009213693951).sum() % 2305843009213693951), int(a.shape[0])
