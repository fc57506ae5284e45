#!/usr/bin/env python
"""print the interesting numbers of bench.py JSON lines (stdin or files)"""
import json, sys
for path in sys.argv[1:] or ["/dev/stdin"]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        b = d.get("breakdown_ms", {})
        print("%s: %.2f ms/step %.1fM sites/s | K1 %.3f (frac %.3f) ev_tot %.3f | gen %.3f count %.3f fb %.3f ptab %.3f pval %.3f assoc %.3f | e2e %.1f ms %s" % (
            path, d["ms_per_step"], d["value"] / 1e6, b.get("ms_events_kernel", 0), (d.get("roofline") or {}).get("frac", 0),
            b.get("ms_events_total", 0), b.get("ms_null_gen", 0), b.get("ms_null_count", 0), b.get("ms_null_fallback", 0),
            b.get("ms_ptable", 0), b.get("ms_pvalues", 0), b.get("ms_assoc_total", 0), d["e2e"].get("ms_per_step", 0),
            json.dumps(d["e2e"].get("breakdown_ms", {}))))
