"""Parity metric of SURVEY.md §8(d), shared by ``bench.py`` and ``tests/`` (pure numpy; touches neither the oracle nor the kernels).

The contract: every output entry compared with the fp64 oracle on identical inputs,

    strict relative error  e = |got - ref| / max(|ref|, 1e-9)  <=  1e-6 (fp64)  /  1e-3 (fp32).

An entry that is a *difference of much larger operands* cannot meet a relative bound in any floating-point format: the defect
``g = x_{k+1} - f(x_k, u_k)`` is O(1e-3) built from O(1) states, so a correctly rounded fp32 evaluation carries an absolute error of
a few ulp(1) ~ 1e-7 and a relative one of 1e-4 .. 1 wherever the defect happens to be small.  Those entries are not hidden behind a
looser tolerance; they are *characterised*: an entry that misses the strict bound is accepted only if its absolute error stays
within ``CANCEL_ULPS`` units of roundoff of the magnitude ``scale`` of the operands it is built from,

    |got - ref| <= CANCEL_ULPS * u(dtype) * scale,     u(f32) = 2^-24, u(f64) = 2^-53,

and the report says how many entries needed that rule and how close the worst one came to it.  ``scale`` is stated per block by the
caller: for ``g`` the largest state magnitude of the batch (the documented cancellation scale), for every other block the block's
own largest entry (sums of products of that size).  ``legacy_rel_err`` is round 1's looser figure, printed beside the strict one.
"""
from __future__ import annotations

import numpy as np

ABS_FLOOR = 1e-9
CANCEL_ULPS = 64.0
UNIT_ROUNDOFF = {"f32": 2.0 ** -24, "f64": 2.0 ** -53}
RTOL = {"f32": 1e-3, "f64": 1e-6}


def strict_rel_err(got, ref) -> np.ndarray:
    """Entry-wise |got - ref| / max(|ref|, 1e-9)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(got - ref) / np.maximum(np.abs(ref), ABS_FLOOR)


def legacy_rel_err(got, ref, scale=None) -> float:
    """Round 1's metric: |got - ref| / (|ref| + 1e-3 * scale), scale = largest |ref| unless given."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    s = float(np.max(np.abs(ref))) if scale is None else float(scale)
    return float(np.max(np.abs(got - ref) / (np.abs(ref) + 1e-3 * s + 1e-300)))


def check_block(got, ref, dtype: str, scale=None) -> dict:
    """Strict gate + cancellation characterisation of one block.  Returns a report; ``ok`` is the verdict."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if got.shape != ref.shape:
        return {"ok": False, "reason": f"shape {got.shape} != {ref.shape}"}
    if ref.size == 0:
        return {"ok": True, "entries": 0, "strict_max": 0.0, "strict_failures": 0, "cancellation_entries": 0,
                "cancellation_worst_ulps": 0.0, "legacy_max": 0.0}
    rtol, u = RTOL[dtype], UNIT_ROUNDOFF[dtype]
    s = float(np.max(np.abs(ref))) if scale is None else float(scale)
    err = np.abs(got - ref)
    strict = err / np.maximum(np.abs(ref), ABS_FLOOR)
    finite = bool(np.all(np.isfinite(got)))
    miss = strict > rtol
    ulps = err / (u * max(s, 1e-300))
    bad = miss & (ulps > CANCEL_ULPS)
    return {
        "ok": finite and not bool(bad.any()),
        "entries": int(ref.size),
        "strict_max_all": float(strict.max()),                       # over every entry, cancellation entries included
        "strict_max": float(strict[~miss].max()) if (~miss).any() else 0.0,  # over the entries that meet the strict bound
        "strict_failures": int(bad.sum()),
        "cancellation_entries": int((miss & ~bad).sum()),
        "cancellation_worst_ulps": float(ulps[miss].max()) if miss.any() else 0.0,
        "scale": s,
        "legacy_max": legacy_rel_err(got, ref, scale),
        "finite": finite,
    }


def merge_reports(reports: dict) -> dict:
    """Fold per-block reports (name -> check_block result) into one summary for a JSON line."""
    ok = all(r.get("ok", False) for r in reports.values())
    entries = sum(r.get("entries", 0) for r in reports.values())
    cancel = sum(r.get("cancellation_entries", 0) for r in reports.values())
    worst_block = max(reports, key=lambda k: reports[k].get("strict_max", 0.0)) if reports else None
    return {
        "ok": ok,
        "metric": "strict: |got-ref| / max(|ref|, 1e-9); entries missing it must lie within 64 units of roundoff of their operand scale",
        "entries": entries,
        "max_rel_err": max((r.get("strict_max", 0.0) for r in reports.values()), default=0.0),
        "max_rel_err_including_cancellation_entries": max((r.get("strict_max_all", 0.0) for r in reports.values()), default=0.0),
        "strict_failures": sum(r.get("strict_failures", 0) for r in reports.values()),
        "cancellation_entries": cancel,
        "cancellation_fraction": cancel / entries if entries else 0.0,
        "cancellation_worst_ulps": max((r.get("cancellation_worst_ulps", 0.0) for r in reports.values()), default=0.0),
        "cancellation_bound_ulps": CANCEL_ULPS,
        "legacy_max_rel_err": max((r.get("legacy_max", 0.0) for r in reports.values()), default=0.0),
        "worst_block": worst_block,
    }


def compare_records(split, got, ref, xp, nX: int, dtype: str, keys=None) -> dict:
    """Block-by-block report for records ``got`` vs ``ref`` (``split`` = Model.split_record).  ``g`` uses the state scale."""
    gb, rb = split(np.asarray(got, dtype=np.float64)), split(np.asarray(ref, dtype=np.float64))
    state_scale = float(np.max(np.abs(np.asarray(xp)[..., :nX])))
    reports = {}
    for key in rb:
        if keys is not None and key not in keys:
            continue
        reports[key] = check_block(gb[key], rb[key], dtype, scale=state_scale if key == "g" else None)
    out = merge_reports(reports)
    out["blocks"] = {k: {kk: r[kk] for kk in ("strict_max", "cancellation_entries", "cancellation_worst_ulps", "legacy_max") if kk in r}
                     for k, r in reports.items()}
    return out
