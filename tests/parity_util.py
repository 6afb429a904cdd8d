"""Comparison core of the GPU parity tests (TEST ONLY).

Tolerances (BASELINE.json north_star): <= 1e-4 relative for Float32, <= 1e-6 for Float64,
norm-wise per tree like the reference's `≈`.  EVERY `complete` tree is compared; none is skipped.
A tree lands in one class:

  strict       norm-wise error <= the north_star tolerance
  loose        above it, but within 30x the tree's own conditioning yardstick — how far the ORACLE's
               result moves under a 1-ulp nudge of X, a precision change, or a 1-ulp move of every
               transcendental result — and that yardstick is small (30 x cond <= CAP = 1e-2)
  elementwise  the norm is dominated by ill-conditioned samples (a pole such as
               x3 / (0.997 - sin(x3 + x2)) hit by a few of 65 536 samples; cos(exp(exp(x))) on part
               of the domain).  The oracle's yardsticks say WHICH elements those are: element j is
               well-conditioned when 30 x (how far that very element moves in the yardsticks) is
               within the north_star tolerance.  Then (1) over the well-conditioned elements the
               device must meet the north_star tolerance norm-wise, and (2) over the others its
               deviation must stay within 30x the oracle's own movement there, norm-wise
  failed       anything else -> the test fails

The caller's asserts bound the classes (at least `min_strict` of the complete trees strict, none
failed).  The distribution (max, p99, class counts) is appended to gpurun_out/parity_stats.jsonl so
that it can be committed under profiles/ (SURVEY.md §7.3-5).
"""
import json
import os

import numpy as np

RTOL = {np.float32: 1e-4, np.float64: 1e-6}
CAP = 1e-2
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b), initial=0.0)), 1e-300)   # avoid overflow inside norm()
    with np.errstate(all="ignore"):
        den = max(np.linalg.norm(b / scale), 1e-300)
        return float(np.linalg.norm(a / scale - b / scale) / den)


def same_nonfinite(a, b):
    return bool((np.isnan(a) == np.isnan(b)).all() and (np.isposinf(a) == np.isposinf(b)).all() and
                (np.isneginf(a) == np.isneginf(b)).all())


def record(label, stats):
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_stats.jsonl"), "a") as f:
            f.write(json.dumps({"label": label, **stats}) + "\n")
    except OSError:
        pass


STRICT, LOOSE, ELEMENTWISE, FAILED = 0, 1, 2, 3
CLASS_NAMES = ("strict", "loose", "elementwise", "failed")


def _norm(v):
    if v.size == 0:
        return 0.0
    m = float(np.max(np.abs(v)))
    if not np.isfinite(m) or m == 0.0:
        return m
    return m * float(np.linalg.norm(v / m))


def part_verdict(dtype, out, ref, yards, mask=None):
    """One output array of one tree (a value row, a gradient block).  yards: the oracle's result
    under the yardstick perturbations.  mask: elements to compare (default: all).
    Returns (class, norm-wise err, norm-wise cond, fraction of well-conditioned elements)."""
    rtol = RTOL[dtype]
    out = np.asarray(out, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    ys = [np.asarray(y, dtype=np.float64).ravel() for y in yards]
    if mask is not None:
        m = np.asarray(mask).ravel()
        out, ref, ys = out[m], ref[m], [y[m] for y in ys]
    if ref.size == 0:
        return STRICT, 0.0, 0.0, 1.0
    err = relerr(out, ref)
    with np.errstate(all="ignore"):
        cond = max((relerr(y, ref) for y in ys), default=0.0)
        if err <= rtol:
            return STRICT, err, cond, 1.0
        if np.isfinite(cond) and 30.0 * cond <= CAP and err <= 30.0 * cond:
            return LOOSE, err, cond, 1.0
        d = np.abs(out - ref)
        sens = np.zeros_like(ref)
        for y in ys:
            sens = np.fmax(sens, np.where(np.isfinite(y), np.abs(y - ref), np.inf))
        well = 30.0 * sens <= rtol * np.abs(ref)
        ill = ~well
        nd_w, nr_w, nd_i, ns_i = _norm(d[well]), _norm(ref[well]), _norm(d[ill]), _norm(sens[ill])
        ok_well = nd_w <= rtol * nr_w
        ok_ill = (not np.isfinite(ns_i)) or nd_i <= 30.0 * ns_i
    detail = {"err_well_conditioned": nd_w / max(nr_w, 1e-300), "dev_ill_conditioned": nd_i, "oracle_movement_ill": ns_i,
              "n": int(ref.size), "n_nonfinite_out": int((~np.isfinite(out)).sum())}
    return (ELEMENTWISE if (ok_well and ok_ill) else FAILED), err, cond, float(well.mean()), detail


def tree_verdict(dtype, parts):
    """parts: [(out, ref, yards[, mask])] of one tree; the tree's class is its worst part's."""
    vs = [part_verdict(dtype, *p) for p in parts]
    k = max(range(len(vs)), key=lambda i: (vs[i][0], vs[i][1] if np.isfinite(vs[i][1]) else np.inf))
    return vs[k]


def check_trees(label, dtype, verdicts, *, min_strict=0.85, ids=None):
    """verdicts: tree_verdict() of every complete tree.  Asserts that none failed and that at
    least `min_strict` of them are strict; records and returns the distribution."""
    cls = np.array([v[0] for v in verdicts], dtype=np.int64)
    errs = np.array([v[1] for v in verdicts], dtype=np.float64)
    n = len(verdicts)
    normwise = errs[(cls == STRICT) | (cls == LOOSE)]
    stats = {
        "dtype": np.dtype(dtype).name, "n_complete": n, "n_strict": int((cls == STRICT).sum()),
        "n_loose": int((cls == LOOSE).sum()), "n_elementwise": int((cls == ELEMENTWISE).sum()),
        "n_failed": int((cls == FAILED).sum()), "rtol": RTOL[dtype],
        "err_max_normwise": float(normwise.max()) if normwise.size else 0.0,
        "err_p99_normwise": float(np.percentile(normwise, 99)) if normwise.size else 0.0,
        "err_median_normwise": float(np.median(normwise)) if normwise.size else 0.0,
        "min_well_conditioned_fraction": float(min((v[3] for v in verdicts), default=1.0)),
    }
    record(label, stats)
    if (cls == FAILED).any():
        k = int(np.nonzero(cls == FAILED)[0][0])
        who = ids[k] if ids is not None else k
        raise AssertionError(f"{label}: tree {who}: norm-wise rel err {errs[k]:.3e} (cond {verdicts[k][2]:.3e}); "
                             f"{verdicts[k][3]:.4f} of its elements are well-conditioned by the oracle's yardsticks and "
                             f"either those miss {RTOL[dtype]:.0e} or the rest exceeds 30x the oracle's own movement: "
                             f"{verdicts[k][4] if len(verdicts[k]) > 4 else ''}; {stats}")
    if n:
        assert stats["n_strict"] >= min_strict * n, \
            f"{label}: only {stats['n_strict']} of {n} complete trees within {RTOL[dtype]:.0e} norm-wise: {stats}"
    return stats
