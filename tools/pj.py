"""Print the interesting fields of a bench.py JSON line (file argument)."""
import json, sys
for line in open(sys.argv[1]):
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    print("value=%.0f e2e=%s frac=%.3f exec_frac=%s ms/step=%.3f clocks=%s" % (
        d["value"], (d.get("e2e") or {}).get("value"), r.get("frac", 0), r.get("executed_frac_of_peak"), d["ms_per_step"], d.get("clocks")))
    print("  ", r.get("per_kernel_ms"))
