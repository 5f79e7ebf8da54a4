"""Prints the per-kernel lines of a bench.py JSON line (developer aid)."""
import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable", e); continue
    r = d["roofline"]
    print(path, "value %.0f q/s  %.1f ms/step" % (d["value"], d["ms_per_step"]))
    for k in r["emitter_pass"]["kernels"]:
        print("   %-20s %8.3f ms  %7.0f GB/s  share %.3f" % (k["kernel"], k["avg_launch_ms"], k["achieved"], k["share_of_step"]))
    print("   pass %.3f ms %.0f GB/s frac %.3f" % (r["emitter_pass"]["ms"], r["emitter_pass"]["achieved"], r["emitter_pass"]["frac"]))
    print("   stages", {k: round(v, 2) for k, v in d["stage_ms_per_step"].items()})
    if "packed" in d:
        p = d["packed"]
        print("   packed %.0f q/s %.1f ms/step" % (p["value"], p["ms_per_step"]), {k: round(v, 2) for k, v in p["stage_ms_per_step"].items()},
              "emit %.0f GB/s" % p["emitter"]["achieved"])
    if "e2e" in d:
        e = d["e2e"]
        print("   e2e %.1f q/s (%s)  dense %.1f" % (e["value"], e.get("format"), e.get("dense", {}).get("value", float("nan"))))
