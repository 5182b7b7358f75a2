"""Pretty-print a bench.py JSON line (per-class and per-layer rows)."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value %.1f %s  ms/step %.1f  e2e %.1f  launches %s  clocks %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), d.get("clocks")))
for k in d["kernels"]:
    print("  %-22s %8.2f ms %5.1f%%  %8.1f %-8s %-6s frac %.3f  (tensor %.3f, hbm %.3f)" % (k["name"], k["ms_per_step"], 100 * k["share"], k.get("achieved", 0), k.get("unit", ""), k.get("bound", ""), k.get("frac", 0), k.get("frac_tensor", 0), k.get("frac_hbm", 0)))
for l in d.get("layers", []):
    print("    %-34s n=%-4g %8.3f ms  %8s TF/s %8s GB/s" % (l["name"], l["launches_per_step"], l["ms_per_step"], l.get("tflops", ""), l.get("gbs", "")))
