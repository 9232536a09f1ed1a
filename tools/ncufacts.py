"""Facts bench.py quotes from an `ncu --page raw --csv` export: python tools/ncufacts.py raw.csv out.json "<kernel description>" [row]
row = which captured launch (default: the last)."""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
data = [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]
d = data[int(sys.argv[4]) if len(sys.argv) > 4 else -1]
num = lambda k: float(d[k].replace(",", ""))
unit = dict(zip(hdr, rows[1]))
def nbytes(k):
    v, u = num(k), unit[k].lower()
    return int(v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u])
rd, wr = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
t = num("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}[unit["gpu__time_duration.sum"].lower()]
out = {"kernel": sys.argv[3], "kernel_name": d["Kernel Name"][:120],
       "source": "ncu --set full --clock-control none (tools/ncu_r02.sh; counters: profiles/r02_ncu_summary.txt, raw: profiles/" + sys.argv[1].split("/")[-1] + ")",
       "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr, "gpu_time_ms_under_ncu": round(t, 4),
       "sm__warps_active_pct": round(num("sm__warps_active.avg.pct_of_peak_sustained_active"), 1),
       "lsu_wavefronts_pct": round(num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"), 1)}
try:
    out["shared_bank_conflict_share"] = round(num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), 3)
except (KeyError, ZeroDivisionError):
    pass
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
