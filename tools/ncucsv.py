"""Print per-kernel metrics of an `ncu --csv --log-file` capture: python tools/ncucsv.py file.csv ..."""
import collections, csv, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    hdr = rows[0]
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        per.setdefault((d["ID"], d["Kernel Name"][:48]), {})[d["Metric Name"]] = (d["Metric Value"], d["Metric Unit"])
    for (i, k), m in per.items():
        print(f, i, k, " | ".join(f"{a.split('__', 1)[1][:42]}={b[0]}{b[1]}" for a, b in m.items()))
