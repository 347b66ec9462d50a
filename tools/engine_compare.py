#!/usr/bin/env python
"""Compare two recordings of tools/engine_trace.py (--record): the reference-CUDA arm and the
b200 arm of the same trace through the same reference engine.

* schedule: with identical ordering scores the scheduler decisions must be IDENTICAL step by step
  (the LTR tick is integer work: bit-exact) - every step's [(request_id, token_chunk_size)].
* tokens: greedy decoding.  The two arms run different attention / norm / rope kernels, so
  logits agree to fp16 round-off, not bit for bit; a greedy token can therefore differ only where
  the reference's own top-2 logits are within round-off of each other.  For every request the
  generated tokens must be identical up to the first such near-tie, and at that position the b200
  token must be one the reference arm itself ranked within `tie_tol` (in log-probability) of its
  choice (needs the reference arm recorded with --logprobs).  After a legitimate tie-break the two
  continuations are different texts and are not compared.
"""
import argparse
import json
import sys


def compare(a: dict, b: dict, tie_tol: float = 0.05) -> dict:
    """a = reference arm, b = b200 arm (dicts as written by engine_trace.py --record)."""
    da, db = a["detail"], b["detail"]
    res = {"schedule_steps": [len(da["schedule"]), len(db["schedule"])]}
    res["schedule_identical"] = da["schedule"] == db["schedule"]
    if not res["schedule_identical"]:
        for i, (x, y) in enumerate(zip(da["schedule"], db["schedule"])):
            if x != y:
                res["schedule_first_diff_step"] = i
                break
    ta, tb = da["tokens"], db["tokens"]
    res["requests"] = [len(ta), len(tb)]
    res["same_finished_set"] = sorted(ta) == sorted(tb)
    lp = da.get("logprobs") or {}
    identical = diverged = unjustified = compared_tokens = 0
    worst_gap = 0.0
    examples = []
    for rid in sorted(set(ta) & set(tb), key=int):
        x, y = ta[rid], tb[rid]
        n = min(len(x), len(y))
        first = next((i for i in range(n) if x[i] != y[i]), None)
        if first is None and len(x) == len(y):
            identical += 1
            compared_tokens += n
            continue
        if first is None:           # same prefix, different length: the stop rule differs
            unjustified += 1
            examples.append({"request": rid, "why": "length", "len": [len(x), len(y)]})
            continue
        diverged += 1
        compared_tokens += first
        row = (lp.get(rid) or [])
        row = row[first] if first < len(row) else None
        if row is None:
            unjustified += 1
            examples.append({"request": rid, "pos": first, "why": "no logprobs recorded"})
            continue
        row = {int(k): v for k, v in row.items()}
        if y[first] not in row:
            unjustified += 1
            examples.append({"request": rid, "pos": first, "why": "b200 token not in the "
                             "reference's top-k", "ref": x[first], "b200": y[first]})
            continue
        gap = row[x[first]] - row[y[first]]
        worst_gap = max(worst_gap, gap)
        if gap > tie_tol:
            unjustified += 1
            examples.append({"request": rid, "pos": first, "why": "gap", "gap": gap})
    res.update(identical_requests=identical, diverged_at_near_tie=diverged - unjustified
               if diverged >= unjustified else 0, unjustified=unjustified,
               tokens_compared=compared_tokens, worst_tie_gap_logprob=worst_gap,
               tie_tol=tie_tol, examples=examples[:8])
    pa, pb = da.get("predicted_scores") or {}, db.get("predicted_scores") or {}
    common = sorted(set(pa) & set(pb), key=int)
    if len(common) > 2:
        import numpy as np
        import scipy.stats
        va, vb = np.array([pa[k] for k in common]), np.array([pb[k] for k in common])
        res["predictor"] = {
            "n": len(common), "max_abs_diff": float(np.abs(va - vb).max()),
            "kendall_tau": float(scipy.stats.kendalltau(va, vb).statistic)}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ref")
    ap.add_argument("b200")
    ap.add_argument("--tie-tol", type=float, default=0.05)
    args = ap.parse_args()
    res = compare(json.load(open(args.ref)), json.load(open(args.b200)), args.tie_tol)
    print(json.dumps(res, indent=1))
    ok = res["schedule_identical"] and res["same_finished_set"] and res["unjustified"] == 0
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
