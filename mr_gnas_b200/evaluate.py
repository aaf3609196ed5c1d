"""Evaluation path of the LP scripts: `predict` / `infer` with the reference's signatures and result dictionaries
(train/mr_lp_train.py:269-347), re-designed around two facts:

  * under model.eval() BatchNorm uses its running statistics, so the full-graph message passing gives the SAME
    entity / relation tables for every evaluation batch -- the reference recomputes them per batch
    (model_lp.py:130-131, ~300 full-graph forwards per epoch at FB15k-237 size); here they are computed once;
  * the filtered rank needs no sort: `mrg_filtered_rank` counts, per query, the unfiltered entities that score
    above the target (ties broken by entity id, i.e. the stable descending order) in one pass over [B, N].
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from ._lib import call, ptr, stream


def filtered_rank(pred, labels, obj):
    """[B] int64 filtered ranks (1 = best) of entity obj[b] among pred[b, :] (see mrg_filtered_rank)."""
    pred, labels = pred.float().contiguous(), labels.float().contiguous()
    obj = obj.long().contiguous()
    B, N = pred.shape
    rank = torch.empty(B, dtype=torch.int32, device=pred.device)
    call("mrg_filtered_rank", ptr(pred), ptr(labels), ptr(obj), B, N, ptr(rank), stream())
    return rank.long()


def embed_once(model, g):
    """Entity / relation tables after all cells in eval mode (identical for every batch of one evaluation)."""
    model.eval()
    with torch.no_grad():
        return model._embed(g)


def predict(val_test_loader, g, model, device, tables=None):
    """Same contract as the reference's predict(): (results, summed BCE test loss); results holds the SUMS
    'mr', 'mrr', 'hits@1/3/10' and 'count' over the loader."""
    with torch.no_grad():
        model.eval()
        all_ent, rel_embed = tables if tables is not None else embed_once(model, g)
        results, test_loss = dict(), []
        for step, (triplets, labels) in enumerate(val_test_loader):
            triplets, labels = triplets.to(device), labels.to(device)
            subj, rel, obj = triplets[:, 0], triplets[:, 1], triplets[:, 2]
            pred = model.score_func(all_ent, all_ent[subj], rel_embed[rel])
            test_loss.append(F.binary_cross_entropy(pred, labels).item())
            ranks = filtered_rank(pred, labels, obj).float()
            results['count'] = torch.numel(ranks) + results.get('count', 0)
            results['mr'] = torch.sum(ranks).item() + results.get('mr', 0)
            results['mrr'] = torch.sum(1.0 / ranks).item() + results.get('mrr', 0)
            for k in [1, 3, 10]:
                results[f'hits@{k}'] = torch.numel(ranks[ranks <= k]) + results.get(f'hits@{k}', 0)
        return results, np.sum(test_loss)


def infer(model, g, tail_loader, head_loader, device):
    """reference: infer() (train/mr_lp_train.py:317-347) without the logging: combined left/right metrics."""
    tables = embed_once(model, g)
    left, left_loss = predict(tail_loader, g, model, device, tables)
    right, right_loss = predict(head_loader, g, model, device, tables)
    assert left['count'] == right['count']
    count = float(left['count'])
    results = {'left_mr': round(left['mr'] / count, 5), 'left_mrr': round(left['mrr'] / count, 5),
               'right_mr': round(right['mr'] / count, 5), 'right_mrr': round(right['mrr'] / count, 5),
               'mr': round((left['mr'] + right['mr']) / (2 * count), 5),
               'mrr': round((left['mrr'] + right['mrr']) / (2 * count), 5)}
    for k in [1, 3, 10]:
        results[f'left_hits@{k}'] = round(left[f'hits@{k}'] / count, 5)
        results[f'right_hits@{k}'] = round(right[f'hits@{k}'] / count, 5)
        results[f'hits@{k}'] = round((results[f'left_hits@{k}'] + results[f'right_hits@{k}']) / 2, 5)
    return results, 0.5 * (left_loss + right_loss)
