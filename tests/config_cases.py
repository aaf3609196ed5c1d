"""Inputs of the BASELINE.json configuration-shaped parity cases (C1 LP train step, C2 NC block step, C3 LP supernet
step), regenerated from seeds exactly as oracle/make_golden.py (gen_config_*) built them for the REAL reference and
verified against the checksums stored in tests/golden/config_c*.pt.  Test infrastructure."""
import os
import types
from collections import namedtuple

import numpy as np
import torch

from oracle import mrg_oracle as O
from oracle.summary import checksum, errors

Genotype = namedtuple("Genotype", "alpha_cell concat_node score_func", defaults=(None,))
Z_SAT = 17.32868                 # fp32 sigmoid(z) rounds to exactly 1.0 above ln(2^25): BCELoss then takes its -100 clamp
JUMP = 100.0 - 16.635532         # ... so the reference's own loss jumps by (100 - 24 ln 2) / (B N) per such element


def lp_args(D):
    return types.SimpleNamespace(feature_dim=D, drop_aggr=0.0, drop_op=0.0, gamma=40, embed_dim=D, conve_hid_drop=0.0,
                                 feat_drop=0.0, num_filt=4, ker_sz=3, k_w=4, k_h=D // 4)


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def c1_inputs(G):
    """-> (triples, subj, rel, labels) of the C1 step: first B items of process()['train'], label smoothing 0.1."""
    d = G["dims"]
    trip = O.synth_kg(d["N"], d["R"], d["T"], seed=0)
    assert checksum(trip) == tuple(G["inputs"]["triples"])
    items = O.process_1n(trip, d["R"])[:d["B"]]
    subj = torch.tensor([it["triple"][0] for it in items])
    rel = torch.tensor([it["triple"][1] for it in items])
    assert torch.equal(subj, G["inputs"]["subj"]) and torch.equal(rel, G["inputs"]["rel"])
    labels = O.smoothed_labels(items, d["N"], 0.1)
    assert max(errors("labels", labels, G["inputs"]["labels"])) == 0.0
    return trip, subj, rel, labels


def c3_inputs(G):
    """-> dict from oracle.sample_search_graph with np.random.seed(0): the sampled 30,000-triple search graph."""
    d = G["dims"]
    trip = O.synth_kg(d["N"], d["R"], d["T"], seed=0)
    assert checksum(trip) == tuple(G["inputs"]["triples"])
    np.random.seed(0)
    s = O.sample_search_graph(trip, d["graph_batch_size"], 0.5, d["R"], d["negative_sample"])
    for key, arr in (("node_id", s["uniq_v"]), ("src", s["src"]), ("dst", s["dst"]), ("etype", s["etype"]),
                     ("samples", s["samples"])):
        assert checksum(arr) == tuple(G["inputs"][key]), key
    assert max(errors("norm", torch.from_numpy(s["norm"]).view(-1, 1), G["inputs"]["norm"])) == 0.0
    return s


def c2_inputs(G):
    d = G["dims"]
    gr = O.synth_nc_graph(d["N"], d["ET"], d["E"], d["C"], 176, seed=0)
    for key in ("src", "dst", "etype"):
        assert checksum(gr[key]) == tuple(G["inputs"][key]), key
    seeds = np.sort(gr["labelled"][:d["B"]])
    assert np.array_equal(seeds, G["inputs"]["seeds"].numpy())
    return gr, seeds


def boundary_count(logits, width):
    """Number of 1-N logits within `width` of the fp32 saturation point of sigmoid (see Z_SAT): these may land on
    either side of the reference loss's discontinuity under any fp32 re-ordering of the GEMM."""
    return int(((logits.double() - Z_SAT).abs() <= width).sum())


def loss_bar(z_full, z_a, z_b, loss_ref, rtol=1e-5, window=2.0):
    """Tolerance for comparing two fp32 evaluations of the reference's BCELoss(sigmoid(z)) at an init where part of
    the probabilities round to exactly 1.0: the loss is DISCONTINUOUS in z at Z_SAT (jump JUMP / numel per element:
    -log(1-p) = 24 ln 2 is replaced by BCELoss's -100 clamp).  `z_a`, `z_b`: the same logits from the two
    evaluations (full tensors or samples at the same positions); their largest disagreement inside
    |z - Z_SAT| < window is the width of the band in which an element may fall on either side; `z_full` (any one
    evaluation, all elements) says how many elements sit in that band.  Bar = rtol |loss| + count * JUMP / numel.
    -> (bar, band width, count)."""
    za, zb = z_a.detach().double().reshape(-1), z_b.detach().double().reshape(-1)
    near = (za - Z_SAT).abs() < window
    width = 2.0 * float((za[near] - zb[near]).abs().max()) if bool(near.any()) else 0.0
    n = boundary_count(z_full, width)
    return rtol * abs(float(loss_ref)) + n * JUMP / z_full.numel(), width, n


def ref_error(summ32, summ64):
    """The fp32 reference's own error against the fp64 evaluation, from the two stored summaries:
    (max |sampled diff| / absmax64, |norm32 - norm64| / norm64)."""
    scale = summ64["absmax"]
    d = float((summ32["vals"].double() - summ64["vals"].double()).abs().max()) if summ32["vals"].numel() else 0.0
    n = abs(summ32["norm"] - summ64["norm"])
    return (d if scale == 0.0 else d / scale), (n if summ64["norm"] == 0.0 else n / summ64["norm"])


def check_vs_truth(name, got, summ32, summ64, tol=1e-5, slack=4.0, report=None):
    """`got` (product or oracle, full tensor) against the fp64 truth summary; bar = max(tol, slack x the fp32
    reference's own error against the same truth) -- separately for sampled values (max-norm) and the 2-norm."""
    e_val, e_nrm = errors(name, got, summ64)
    r_val, r_nrm = ref_error(summ32, summ64)
    ok = e_val <= max(tol, slack * r_val) and e_nrm <= max(tol, slack * r_nrm)
    if report is not None:          # collect: the caller asserts once over all tensors (assert_report)
        report.append((e_val, r_val, e_nrm, r_nrm, name, ok, summ64["absmax"]))
        return e_val
    assert e_val <= max(tol, slack * r_val), f"{name}: sampled rel err {e_val:.2e} (reference's own {r_val:.2e})"
    assert e_nrm <= max(tol, slack * r_nrm), f"{name}: 2-norm rel err {e_nrm:.2e} (reference's own {r_nrm:.2e})"
    return e_val


def assert_report(tag, rep, tol=1e-5, slack=4.0, max_fraction=0.10, hard=32.0):
    """Verdict over a check_vs_truth(report=...) collection.

    Per tensor the bar is max(tol, slack x the real fp32 reference's own error against the fp64 truth).  The
    reference's error is ONE draw of rounding noise, and so is ours: for two independent, equally accurate
    evaluations the ratio |e_ours| / |e_ref| exceeds 4 with probability (2/pi) atan(1/4) = 16 % (ratio of two
    centred normals is Cauchy), so over hundreds of tensors some MUST cross a fixed multiple without being less
    accurate.  Hence: at most `max_fraction` of the tensors beyond their bar, and none beyond max(tol, hard x its
    reference error) -- "as accurate as the reference's own fp32 path", stated so that it is falsifiable."""
    live = [r for r in rep if r[1] < 1.0] or rep     # (a tensor whose true value is 0 has no relative error)
    worst = max(live)
    print(f"{tag}: {len(rep)} tensors; worst sampled err vs fp64 truth: ours %.2e, real reference's own %.2e "
          f"(2-norm %.2e / %.2e) at %s" % worst[:5])
    ratios = sorted(max(r[0] / max(r[1], tol / slack), r[2] / max(r[3], tol / slack)) for r in rep)
    print(f"  error ratio ours / max(reference's own, {tol / slack:.1e}): median {ratios[len(ratios) // 2]:.2f}, "
          f"90th percentile {ratios[int(0.9 * (len(ratios) - 1))]:.2f}, max {ratios[-1]:.2f}")
    bad = [r for r in rep if not r[5]]
    for r in sorted(bad, reverse=True):
        print("  beyond 4x   %-70s ours %.2e ref %.2e | 2-norm ours %.2e ref %.2e | absmax %.2e" %
              (r[4], r[0], r[1], r[2], r[3], r[6]))
    floor = tol / slack
    way_off = [r for r in rep if r[0] > hard * max(r[1], floor) or r[2] > hard * max(r[3], floor)]
    assert not way_off, f"{tag}: {[r[4] for r in way_off]} exceed max({tol}, {hard} x the reference's own fp32 error)"
    assert len(bad) <= max_fraction * len(rep), \
        f"{tag}: {len(bad)} of {len(rep)} tensors exceed max({tol}, {slack} x the reference's own fp32 error)"
