"""Search-script graph pipeline on the device: the per-step host work of
search/mr_lp_search.py:188-212 -- utils/utils_rgcn.py:79-118 (`generate_sampled_graph_and_labels`: uniform edge
sample, node relabelling, negative sampling, graph split), :129-157 (`build_graph_from_triplets`: both directions,
Python ``sorted(zip(rel, dst, src))``), :120-127 (`comp_deg_norm`) and mr_lp_search.py:30-36
(`node_norm_to_edge_norm`) -- as device-side tensor ops + the K0 graph build (mrg_graph_build), so that a search
step never leaves the GPU: at graph_batch_size 30,000 the reference spends ~0.5 s per step in numpy / Python
sorting here, more than the whole supernet forward+backward takes on the B200.

`sample_search_graph` is the device-native entry point (everything stays a CUDA tensor).  The reference-named
functions below keep the reference's signatures and return types (numpy arrays + a graph object) so that
search/mr_lp_search.py runs unchanged on top of them.

Random numbers: the reference draws from numpy's global generator on the host; the device path draws from a
torch.Generator (or takes the draws explicitly through `draws=` -- how the tests pin it bit-exactly to the
reference pipeline: same chosen triples, same corruptions, same graph half => identical arrays)."""
import numpy as np
import torch

from .graph import MRGraph


def _draws(T, sample_size, split, n_neg, device, generator):
    """The four random draws of the pipeline, in the reference's order (:87, :198-199, :107-108)."""
    g = generator
    edges = torch.randperm(T, device=device, generator=g)[:sample_size]
    values_u = torch.rand(n_neg, device=device, generator=g)          # scaled to [0, n_nodes) once n is known
    choices = torch.rand(n_neg, device=device, generator=g)
    keep = torch.randperm(sample_size, device=device, generator=g)[:split]
    return {"edges": edges, "values_u": values_u, "choices": choices, "keep": keep}


def sample_search_graph(triplets, sample_size, split_size, num_rels, negative_rate, device="cuda", generator=None,
                        draws=None):
    """Device form of generate_sampled_graph_and_labels(sampler="uniform").

    triplets [T, 3] (s, r, o) int64 (any device; moved once).  Returns a dict of device tensors:
      g (MRGraph of the graph half, edges ordered by (rel, dst, src), norm = n_norm[dst] * n_norm[src] as [E, 1]),
      uniq_v [n] global ids of the sampled nodes, src / dst / etype [E] (relabelled, both directions),
      node_norm [n], samples [(1 + negative_rate) * sample_size, 3], labels (float32).
    draws: optional dict(edges, values, choices, keep) of explicit random draws (numpy or tensors): `edges` indices
    into triplets, `values` corrupting entity ids, `choices` uniforms, `keep` indices of the graph half."""
    dev = torch.device(device)
    trip = torch.as_tensor(triplets).to(dev)
    T = trip.shape[0]
    split = int(sample_size * split_size)
    n_neg = sample_size * negative_rate
    d = draws if draws is not None else _draws(T, sample_size, split, n_neg, dev, generator)
    as_t = lambda a: torch.as_tensor(a).to(dev)
    picked = trip[as_t(d["edges"]).long()]
    # relabel nodes to consecutive ids: np.unique((src, dst), return_inverse=True) == sorted unique + inverse
    uniq_v, inv = torch.unique(torch.cat([picked[:, 0], picked[:, 2]]), sorted=True, return_inverse=True)
    n = int(uniq_v.numel())
    s, o = inv[:sample_size], inv[sample_size:]
    r = picked[:, 1]
    pos = torch.stack([s, r, o], 1)
    # negative_sampling (:191-204): corrupt the subject where u > 0.5, the object otherwise
    neg = pos.repeat(negative_rate, 1)
    if "values" in d:
        values = as_t(d["values"]).long()
    else:
        values = (d["values_u"] * n).long().clamp_(max=n - 1)
    choices = as_t(d["choices"])
    subj = choices > 0.5
    neg[:, 0] = torch.where(subj, values, neg[:, 0])
    neg[:, 2] = torch.where(subj, neg[:, 2], values)
    samples = torch.cat([pos, neg])
    labels = torch.zeros(sample_size * (negative_rate + 1), dtype=torch.float32, device=dev)
    labels[:sample_size] = 1
    keep = as_t(d["keep"]).long()
    out = build_graph_from_triplets_device(n, num_rels, s[keep], r[keep], o[keep], dev)
    out.update(uniq_v=uniq_v, samples=samples, labels=labels)
    return out


def build_graph_from_triplets_device(num_nodes, num_rels, s, r, o, device="cuda"):
    """utils_rgcn.build_graph_from_triplets (:129-157) on the device: edges [s->o | o->s], relation ids [r | r+R],
    ordered by (rel, dst, src) -- one radix sort of a packed 64-bit key instead of sorted(zip(...)) --, then the K0
    build (dst-CSR, src-CSC, relation segments, degree norms)."""
    dev = torch.device(device)
    src, dst = torch.cat([s, o]), torch.cat([o, s])
    rel = torch.cat([r, r + num_rels])
    n = max(int(num_nodes), 1)
    key = (rel * n + dst) * n + src            # < (2R) * n^2: fits int64 for every configuration in BASELINE.json
    order = torch.argsort(key, stable=True)
    src, dst, rel = src[order].contiguous(), dst[order].contiguous(), rel[order].contiguous()
    g = MRGraph.from_edges(src, dst, rel, num_nodes, 2 * num_rels + 1, device=dev)
    g.edata['norm'] = g.edge_norm.view(-1, 1)          # node_norm_to_edge_norm leaves [E, 1] (mr_lp_search.py:30-36)
    return {"g": g, "num_nodes": num_nodes, "src": src, "dst": dst, "etype": rel, "node_norm": g.n_norm,
            "norm": g.edge_norm}


# ------------------------------------------------------------------ reference-named wrappers (numpy in / numpy out)
def get_adj_and_degrees(num_nodes, triplets):
    """utils_rgcn.py:18-30.  The uniform sampler never reads the adjacency lists; degrees are returned for API
    compatibility (the "neighbor" sampler -- a sequential random walk -- is host-side in the reference and is not
    part of the device path)."""
    t = np.asarray(triplets)
    deg = np.bincount(np.concatenate([t[:, 0], t[:, 2]]), minlength=num_nodes)
    return None, deg


def generate_sampled_graph_and_labels(triplets, sample_size, split_size, num_rels, adj_list, degrees, negative_rate,
                                      sampler="uniform", device="cuda", generator=None):
    """Same signature and return tuple as utils_rgcn.py:79-118: (g, uniq_v, src, rel, node_norm, samples, labels),
    arrays as numpy (the script calls torch.from_numpy on them, mr_lp_search.py:198-203)."""
    if sampler != "uniform":
        raise ValueError("the device pipeline implements the 'uniform' edge sampler (the reference default)")
    d = sample_search_graph(triplets, sample_size, split_size, num_rels, negative_rate, device, generator)
    npy = lambda t: t.detach().cpu().numpy()
    return (d["g"], npy(d["uniq_v"]), npy(d["src"]), npy(d["etype"]), npy(d["node_norm"]), npy(d["samples"]),
            npy(d["labels"]))


def build_graph_from_triplets(num_nodes, num_rels, triplets, device="cuda"):
    """utils_rgcn.py:129-157: (g, src, rel, node_norm) with numpy arrays; triplets = (src, rel, dst) arrays."""
    s, r, o = (torch.as_tensor(np.asarray(a)).to(device).long() for a in triplets)
    d = build_graph_from_triplets_device(num_nodes, num_rels, s, r, o, device)
    return d["g"], d["src"].cpu().numpy(), d["etype"].cpu().numpy(), d["node_norm"].cpu().numpy()


build_graph_from_triplets_ori = build_graph_from_triplets


def build_test_graph(num_nodes, num_rels, edges, device="cuda"):
    """utils_rgcn.py:186-189."""
    e = np.asarray(edges)
    return build_graph_from_triplets(num_nodes, num_rels, (e[:, 0], e[:, 1], e[:, 2]), device)


def comp_deg_norm(g):
    """utils_rgcn.py:120-127: in_deg ** -0.5 with inf -> 0 (already computed by the K0 build)."""
    return g.n_norm.detach().cpu().numpy()


def negative_sampling(pos_samples, num_entity, negative_rate, device="cuda", generator=None):
    """utils_rgcn.py:191-204 on the device; numpy in / numpy out."""
    dev = torch.device(device)
    pos = torch.as_tensor(np.asarray(pos_samples)).to(dev)
    nb = pos.shape[0]
    neg = pos.repeat(negative_rate, 1)
    values = torch.randint(num_entity, (nb * negative_rate,), device=dev, generator=generator)
    choices = torch.rand(nb * negative_rate, device=dev, generator=generator)
    subj = choices > 0.5
    neg[:, 0] = torch.where(subj, values, neg[:, 0])
    neg[:, 2] = torch.where(subj, neg[:, 2], values)
    labels = np.zeros(nb * (negative_rate + 1), dtype=np.float32)
    labels[:nb] = 1
    return torch.cat([pos, neg]).cpu().numpy(), labels
