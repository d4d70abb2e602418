import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    print("value %.1f Mrays/s  ms/step %.2f  e2e %.1f  kernel_ms %s  frac %.3f  launches %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v, 2) for k, v in (r.get("kernel_ms") or {}).items()},
        r.get("frac", 0), d.get("gpu_launches")))
    if r:
        print("   per_query(ref schedule)", r.get("per_query"), "actual", r.get("per_query_actual"), "cpu", d.get("cpu_baseline"))
