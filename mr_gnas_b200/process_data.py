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
