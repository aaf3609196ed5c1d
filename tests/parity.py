"""Parity metrics shared by the GPU tests (the north_star's "within 1e-5 relative in fp32").

Every float comparison is RELATIVE TO THE REFERENCE TENSOR'S OWN SCALE -- there is no absolute floor, so a
gradient whose entries are all ~1e-6 must still agree to ~1e-11:

  rel_err(a, b)  = max|a - b| / max|b|          (max-norm relative error)
  norm_err(a, b) = ||a - b||_2 / ||b||_2        (norm-wise relative error)

and `check` asserts BOTH <= tol (default RTOL = 1e-5).  An element-wise bound |a-b| <= rtol*|b| with no scale
term cannot hold for fp32 sums that nearly cancel (the element's own magnitude says nothing about the size of
the terms that were added), so the element-wise form used is |a - b| <= RTOL*|b| + RTOL*max|b|, which the
max-norm bound implies.  When the reference tensor is exactly zero the error is absolute (max|a|).

`truth` (optional): an fp64 evaluation of the same quantity.  fp32 CPU and fp32 GPU sum in different orders, so
against fp64 the bar is max(tol, 4 x the fp32 CPU reference's own error vs fp64): the product may not be further
from the truth than a small multiple of what the reference's fp32 path itself achieves."""
import torch

RTOL = 1e-5


def _d(t):
    return t.detach().double().cpu()


def rel_err(a, b):
    a, b = _d(a), _d(b)
    if a.numel() == 0:
        return 0.0
    scale = float(b.abs().max())
    diff = float((a - b).abs().max())
    return diff if scale == 0.0 else diff / scale


def norm_err(a, b):
    a, b = _d(a), _d(b)
    if a.numel() == 0:
        return 0.0
    scale = float(b.norm())
    diff = float((a - b).norm())
    return diff if scale == 0.0 else diff / scale


def bar(ref32=None, truth=None, tol=RTOL):
    """Tolerance against `truth` given the fp32 reference's own distance from it."""
    if truth is None or ref32 is None:
        return tol
    return max(tol, 4.0 * rel_err(ref32, truth))


def check(name, a, b, tol=RTOL, truth=None):
    """a: product result, b: fp32 reference (oracle / golden), truth: optional fp64 evaluation."""
    assert tuple(a.shape) == tuple(b.shape), f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    if truth is not None:
        t = bar(b, truth, tol)
        e, n = rel_err(a, truth), norm_err(a, truth)
    else:
        t = tol
        e, n = rel_err(a, b), norm_err(a, b)
    assert e <= t and n <= t, f"{name}: max-norm rel err {e:.3e}, 2-norm rel err {n:.3e} > {t:.1e}"
    return e


def grad_errors(named_a, named_b, negligible=1e-5):
    """Per-tensor rel_err of gradient dict `named_a` (product) against `named_b` (reference), as a list of
    (err, name).  A parameter whose TRUE gradient is exactly zero -- the bias of a Linear that feeds a BatchNorm
    (linear_e.bias, concat.bias, the NC OpModule.linear.bias): a shift the normalisation removes -- holds only
    rounding noise on both sides, ~1e-7 of the network's gradient scale, and has no relative error to speak of.
    Such tensors (reference max <= `negligible` x the largest gradient max of the model) are required to be
    negligible in the product too and are reported with err 0."""
    top = max(float(b.detach().abs().max()) for b in named_b.values() if b is not None)
    out = []
    for k, b in named_b.items():
        if b is None:
            continue
        a = named_a[k]
        if float(b.detach().abs().max()) <= negligible * top:
            assert float(a.detach().abs().max()) <= 10 * negligible * top, f"{k}: reference gradient is ~0, ours is not"
            out.append((0.0, k))
        else:
            out.append((rel_err(a, b), k))
    return out


def check_grads(tag, ours, ref32, truth64, tol=RTOL, slack=4.0, max_fraction=0.10, hard=32.0, verbose=True):
    """Gradients (dicts name -> tensor) of the product against the fp64 evaluation `truth64` of the same reference
    modules, knowing the real fp32 reference's own result `ref32`:
      * a tensor whose TRUE gradient is zero (|truth| <= 1e-9 x the model's largest gradient: the bias of a Linear
        feeding a BatchNorm, CompGCN's loop relation under `sub`, ...) must be negligible in the product too;
      * otherwise err = max|ours - truth| / max|truth| must be <= max(tol, slack x the reference's own error) --
        for all but `max_fraction` of the tensors (the ratio of two independent rounding-error draws exceeds 4
        with probability 16 %), and <= hard x max(reference's error, tol / slack) for every tensor.
    Returns the list of (err, ref_err, name)."""
    keys = [k for k, t in truth64.items() if t is not None]
    top = max(float(truth64[k].detach().abs().max()) for k in keys)
    rows, bad = [], []
    for k in keys:
        t = truth64[k].detach().double().cpu()
        if float(t.abs().max()) <= 1e-9 * top:
            assert float(ours[k].detach().abs().max()) <= 1e-4 * top, f"{tag} {k}: true gradient is 0, ours is not"
            continue
        e, r = rel_err(ours[k], t), rel_err(ref32[k], t)
        rows.append((e, r, k))
        assert e <= hard * max(r, tol / slack), f"{tag} {k}: rel err {e:.2e}, the reference's own {r:.2e}"
        if e > max(tol, slack * r):
            bad.append((e, r, k))
    if verbose and rows:
        ratios = sorted(e / max(r, tol / slack) for e, r, _ in rows)
        print(f"{tag}: {len(rows)} gradient tensors vs fp64 truth; worst ours %.2e (reference's own %.2e) at %s; "
              f"ratio ours/reference median {ratios[len(ratios) // 2]:.2f} max {ratios[-1]:.2f}" % max(rows))
    assert len(bad) <= max(1, int(max_fraction * len(rows))), f"{tag}: beyond {slack}x the reference's own error: {bad}"
    return rows


def check_grads_pair(tag, named_a, named_b, tol=2e-5, max_fraction=0.10, hard=2e-3):
    """Two fp32 evaluations of the SAME step that differ only in summation order (e.g. destination-partitioned vs
    single GPU) -- no fp64 truth at hand.  Per tensor max|a-b| / max|b|: well-conditioned tensors agree to `tol`;
    gradients that are small differences of large sums (the W.bias of a sparse gate: the real fp32 reference is
    itself up to 5e-4 from its fp64 value there, see tests/golden/network_nc.pt) move with the order of the sum.
    So: at most `max_fraction` of the tensors beyond `tol`, none beyond `hard`; zero-true-gradient tensors are
    handled as in grad_errors."""
    errs = grad_errors(named_a, named_b)
    beyond = [(e, k) for e, k in errs if e > tol]
    worst = max(errs)
    print(f"{tag}: {len(errs)} gradient tensors, worst rel err {worst[0]:.2e} at {worst[1]}, {len(beyond)} beyond {tol:.0e}")
    assert worst[0] <= hard, (tag, worst)
    assert len(beyond) <= max(1, int(max_fraction * len(errs))), (tag, sorted(beyond, reverse=True))
