"""Print the headline numbers of bench JSON lines: python tools/showbench.py file.json ..."""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f, "N=%d value %.2f G/s  ms %.3f | e2e %.2f G/s ms %.3f | roof %s %.3f" % (d["n_gpus"], d["value"] / 1e9, d["ms_per_step"],
              d["e2e"]["value"] / 1e9, d["e2e"].get("ms_per_step", 0), r.get("kernel"), r.get("frac", 0)))
        print("   ", {k: round(v["ms_per_step"], 3) for k, v in d.get("kernels", {}).items()})
        if d.get("probe"): print("    probe", d["probe"])
        if d.get("cpu_baseline"): print("    cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:80])
        print("    clocks", d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
