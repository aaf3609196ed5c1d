"""Host-side triple preprocessing with the reference's semantics (utils/process_data.py:4-31,
utils/data_set.py:6-59): 1-N training items {(s, r) -> objects} including inverse relations,
dense multi-hot labels with label smoothing.  Also the place where the training graph is
rebuilt as destination-sorted CSR + (relation, direction) segments (north_star): see
``build_graph`` -> mr_gnas_b200.graph.MRGraph."""
from collections import defaultdict as ddict

import numpy as np
import torch

from .graph import MRGraph


def process(dataset, num_rel):
    """reference: utils/process_data.py:4-31 (same keys, same item order)."""
    sr2o = ddict(set)
    for subj, rel, obj in dataset['train']:
        sr2o[(subj, rel)].add(obj)
        sr2o[(obj, rel + num_rel)].add(subj)
    sr2o_train = {k: list(v) for k, v in sr2o.items()}
    for split in ['valid', 'test', 'train']:
        for subj, rel, obj in dataset[split]:
            sr2o[(subj, rel)].add(obj)
            sr2o[(obj, rel + num_rel)].add(subj)
    sr2o_all = {k: list(v) for k, v in sr2o.items()}
    triplets = ddict(list)
    for (subj, rel), obj in sr2o_train.items():
        triplets['train'].append({'triple': (subj, rel, -1), 'label': sr2o_train[(subj, rel)]})
    for split in ['valid', 'test', 'train']:
        for subj, rel, obj in dataset[split]:
            triplets[f"{split}_tail"].append({'triple': (subj, rel, obj), 'label': sr2o_all[(subj, rel)]})
            triplets[f"{split}_head"].append(
                {'triple': (obj, rel + num_rel, subj), 'label': sr2o_all[(obj, rel + num_rel)]})
    return dict(triplets)


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
