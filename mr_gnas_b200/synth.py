"""Synthetic knowledge-graph generators of the benchmark shapes (no datasets / network here).
SURVEY.md section 8d: Zipf(0.8) heads/tails over a random entity permutation, Zipf(1.0)
relations, duplicates allowed (the reference de-duplicates only inside sr2o sets)."""
import numpy as np

CONFIGS = {
    # name: (entities, base relations, train triples, feature dim)
    "c1_fb15k237": (14541, 237, 272115, 200),
    "c3_wn18rr": (40943, 11, 86835, 200),
    "tiny": (500, 7, 4000, 64),
    # BASELINE configs[4] (C5: 10 M entities, 1 k relations, 200 M directed edges = 100 M triples, D = 256) scaled down
    # in entities and triples alike (same relation count and feature dim); the full size needs the entity table and
    # its optimiser state sharded
    "c5_64th": (156_250, 1000, 1_562_500, 256),
    "c5_eighth": (1_250_000, 1000, 12_500_000, 256),
}


def synth_kg(num_ent, num_rels, num_triples, seed=0, alpha_ent=0.8, alpha_rel=1.0):
    rng = np.random.RandomState(seed)
    perm = rng.permutation(num_ent)

    def zipf(n, size, alpha):
        p = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), alpha)
        p /= p.sum()
        return rng.choice(n, size=size, p=p)

    s = perm[zipf(num_ent, num_triples, alpha_ent)]
    o = perm[rng.permutation(num_ent)[zipf(num_ent, num_triples, alpha_ent)]]
    r = zipf(num_rels, num_triples, alpha_rel)
    return np.stack([s, r, o], 1).astype(np.int64)
