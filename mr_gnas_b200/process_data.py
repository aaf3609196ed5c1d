"""Host-side triple preprocessing with the reference's semantics (utils/process_data.py:4-31,
utils/data_set.py:6-59): 1-N training items {(s, r) -> objects} including inverse relations,
dense multi-hot labels with label smoothing.  Also the place where the training graph is
rebuilt as destination-sorted CSR + (relation, direction) segments (north_star): see
``build_graph`` -> mr_gnas_b200.graph.MRGraph."""
import numpy as np
import torch

from .graph import MRGraph


def _queries(triples, num_rel):
    """[T,3] (s, r, o) -> the 2T 1-N queries in the order the reference visits them: (s, r)->o and (o, r+R)->s
    interleaved per triple.  Returns (key [2T] = subj * 2R + rel, answer [2T])."""
    t = np.asarray(triples, dtype=np.int64).reshape(-1, 3)
    subj = np.stack([t[:, 0], t[:, 2]], 1).reshape(-1)
    rel = np.stack([t[:, 1], t[:, 1] + num_rel], 1).reshape(-1)
    return subj * (2 * num_rel) + rel, np.stack([t[:, 2], t[:, 0]], 1).reshape(-1)


class QueryCSR:
    """The (subject, relation) -> {objects} map as one sorted-unique CSR (the form the device label kernel and
    make_batch_sparse want): `keys` ascending, `ptr` [K+1], `obj` ascending inside a key, `first` = position of the
    key's first occurrence in the visiting order (the reference's dict insertion order)."""

    def __init__(self, key, ans):
        order = np.lexsort((ans, key))
        k_s, a_s = key[order], ans[order]
        keep = np.ones(k_s.shape[0], dtype=bool)
        keep[1:] = (k_s[1:] != k_s[:-1]) | (a_s[1:] != a_s[:-1])
        k_u, self.obj = k_s[keep], a_s[keep]
        self.keys, start = np.unique(k_u, return_index=True)
        self.ptr = np.append(start, k_u.shape[0]).astype(np.int64)
        _, self.first = np.unique(key, return_index=True)
        self._lists = None

    def lookup(self, key):
        """rows of `key` in this CSR (every key must be present)."""
        return np.searchsorted(self.keys, key)

    def labels(self, row):
        """object list of CSR row `row` (one shared list per row, like the reference's shared sr2o lists)."""
        if self._lists is None:
            flat, p = self.obj.tolist(), self.ptr.tolist()
            self._lists = [flat[a:b] for a, b in zip(p[:-1], p[1:])]
        return self._lists[row]


def process(dataset, num_rel):
    """Same result dictionary as the reference's utils/process_data.py:4-31 -- 'train' holds one
    {'triple': (s, r, -1), 'label': objects seen in train} item per distinct query in first-seen order,
    '{split}_tail' / '{split}_head' one item per triple with the objects seen in ANY split -- built from two
    sorted-unique CSRs instead of per-triple Python set updates.  Label lists come out ascending (the reference's
    `list(set)` order is arbitrary; every consumer scatters them into a multi-hot row)."""
    R2 = 2 * num_rel
    splits = {s: np.asarray(dataset[s], dtype=np.int64).reshape(-1, 3) for s in ('train', 'valid', 'test')}
    k_tr, a_tr = _queries(splits['train'], num_rel)
    train = QueryCSR(k_tr, a_tr)
    k_all, a_all = _queries(np.concatenate([splits['valid'], splits['test'], splits['train']]), num_rel)
    every = QueryCSR(k_all, a_all)
    seen = np.argsort(train.first, kind='stable')
    out = {'train': [{'triple': (s, r, -1), 'label': train.labels(j)} for s, r, j in
                     zip((train.keys[seen] // R2).tolist(), (train.keys[seen] % R2).tolist(), seen.tolist())]}
    for split in ('valid', 'test', 'train'):
        t = splits[split]
        key, _ = _queries(t, num_rel)
        rows = every.lookup(key).reshape(-1, 2)
        tails, heads = [], []
        for (s, r, o), (jt, jh) in zip(t.tolist(), rows.tolist()):
            tails.append({'triple': (s, r, o), 'label': every.labels(jt)})
            heads.append({'triple': (o, r + num_rel, s), 'label': every.labels(jh)})
        if tails:
            out[f"{split}_tail"], out[f"{split}_head"] = tails, heads
    return out


def train_items(triples, num_rel, select=None):
    """process(...)['train'] (or only its items `select`, positions in that list) without building the evaluation
    splits -- what a training loop that draws query batches needs.  -> (items, number of distinct queries)."""
    R2 = 2 * num_rel
    key, ans = _queries(triples, num_rel)
    csr = QueryCSR(key, ans)
    seen = np.argsort(csr.first, kind='stable')
    total = seen.shape[0]
    if select is not None:
        seen = seen[np.asarray(select, dtype=np.int64)]
    return [{'triple': (int(csr.keys[j] // R2), int(csr.keys[j] % R2), -1),
             'label': csr.obj[csr.ptr[j]:csr.ptr[j + 1]].tolist()} for j in seen.tolist()], total


def build_graph(num_ent, data, num_rels, device="cuda"):
    """reference: train/mr_lp_train.py:77-89, rebuilt on the device as dst-sorted CSR."""
    return MRGraph.from_triples(num_ent, np.asarray(data), num_rels, device=device)


def make_batch(items, num_ent, lbl_smooth=0.0, pin=False):
    """Collated (triplets [B,3] int64, labels [B,N] fp32) as TrainDataset + default collate produce
    (utils/data_set.py:17-33)."""
    B = len(items)
    trip = torch.tensor([list(it['triple']) for it in items], dtype=torch.long)
    y = torch.zeros(B, num_ent, dtype=torch.float32)
    for i, it in enumerate(items):
        y[i, torch.as_tensor(np.int32(it['label']), dtype=torch.long)] = 1.0
    if lbl_smooth != 0.0:
        y = (1.0 - lbl_smooth) * y + (1.0 / num_ent)
    if pin:
        trip, y = trip.pin_memory(), y.pin_memory()
    return trip, y


def make_batch_sparse(items, pin=False):
    """The same batch with the labels in their sparse form: (triplets [B,3] int64, ptr [B+1] int32, idx [nnz]
    int32) -- the object lists of process() (utils/process_data.py:19) as one CSR.  `labels_on_device` expands
    them on the GPU into exactly the rows TrainDataset would have produced (utils/data_set.py:17-33)."""
    trip = torch.tensor([list(it['triple']) for it in items], dtype=torch.long)
    counts = [len(it['label']) for it in items]
    ptr = torch.zeros(len(items) + 1, dtype=torch.int32)
    ptr[1:] = torch.cumsum(torch.tensor(counts, dtype=torch.int64), 0).to(torch.int32)
    idx = torch.tensor([o for it in items for o in it['label']], dtype=torch.int32)
    if pin:
        trip, ptr, idx = trip.pin_memory(), ptr.pin_memory(), idx.pin_memory()
    return trip, ptr, idx


def label_values(num_ent, lbl_smooth):
    """(neg, pos) fp32 label values as torch computes (1.0 - ls) * y + 1.0 / N on an fp32 multi-hot y."""
    if lbl_smooth == 0.0:
        return 0.0, 1.0
    inv = np.float32(1.0 / num_ent)
    return float(inv), float(np.float32(np.float32(1.0 - lbl_smooth) * np.float32(1.0)) + inv)


def labels_on_device(ptr, idx, num_ent, lbl_smooth=0.0, col_lo=0, col_hi=None, out=None):
    """Dense [B, col_hi - col_lo] fp32 labels written on the device from device-resident (ptr, idx)."""
    from . import _lib
    col_hi = num_ent if col_hi is None else col_hi
    B = ptr.numel() - 1
    if out is None:
        out = torch.empty(B, col_hi - col_lo, dtype=torch.float32, device=ptr.device)
    neg, pos = label_values(num_ent, lbl_smooth)
    _lib.call("mrg_labels_from_csr", _lib.ptr(ptr), _lib.ptr(idx), B, col_lo, col_hi, neg, pos, _lib.ptr(out),
              _lib.stream())
    return out
