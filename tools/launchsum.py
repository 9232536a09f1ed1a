"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv --log-file` launch list:
python tools/launchsum.py launches.csv "<command that was profiled>" > summary.txt"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
tot = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(hdr, r))
    if d["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(d["Metric Unit"], 1e-6)
    name = re.sub(r"\(.*$", "", d["Kernel Name"]).strip()
    t = tot.setdefault(name, [0.0, 0])
    t[0] += v; t[1] += 1
total = sum(v[0] for v in tot.values())
n = sum(v[1] for v in tot.values())
print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# (per-launch times are cold-cache and serialised: shares, not absolutes)")
print(f"# {n} launches, {total:.2f} ms of device time\n")
for name, (ms, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * ms / total:6.2f} %  {ms:9.3f} ms  {c:4d} launches  {1e3 * ms / c:9.1f} us each  {name}")
