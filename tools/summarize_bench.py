#!/usr/bin/env python3
"""one line per bench JSON on stdin"""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    k = r.get("kernel_ms") or {}
    print("%.1f Mrays/s  %.2f ms/step  e2e %.1f | trace %.1f shadow %.1f shade %.1f (ms, profiled leg) | frac %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], k.get("ms_trace", 0), k.get("ms_shadow", 0), k.get("ms_shade", 0),
        ("%.3f" % r["frac"]) if r else "-"))
