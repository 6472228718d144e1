"""Condenses gpurun_out/parity_log.jsonl (written by tests/util.py::check_close on the GPU box) into a table:
one line per check with the achieved max|diff|, max|ref| and the fraction of elements outside the unscaled
north-star tolerance (rtol 1e-3 / atol 1e-5)."""
import json, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_log.jsonl"
rows = [json.loads(l) for l in open(path) if l.strip()]
print(f"# {len(rows)} parity checks (tests/util.py::check_close), source {path}")
print(f"{'test':70s} {'what':60s} {'max|diff|':>10s} {'max|ref|':>10s} {'strict_viol':>11s} {'allowed':>8s} {'rtol':>8s}")
for r in rows:
    print(f"{r['test'][-70:]:70s} {r['what'][:60]:60s} {r['max_diff']:10.3e} {r['max_ref']:10.3e} {r['strict_viol_frac']:11.2e} {r['max_strict_viol']:8.1e} {r['rtol']:8.1e}")
