"""Tuning run: per-kernel CUDA-event times of make.kmer.hash for sort-pass configurations.
usage: python tools/sortbench.py [L] [k] [rank:shape:rb:bits,...] [workload c2|c3]
  rank 0 bitmap / 3 one atomic; shape 0 256x24x2, 1 256x28x2, 2 256x20x3 (8-bit only); rb digit bits; bits = hash bits (0 = auto)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import kmer_hasher_b200 as kh
from kmer_hasher_b200 import _lib, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
specs = sys.argv[3].split(",") if len(sys.argv) > 3 else ["3:0:8:32", "3:0:8:40", "3:0:9:36", "3:0:10:40", "3:1:8:32", "3:1:9:36", "3:2:8:32", "0:0:8:32", "0:0:9:36"]
wl = sys.argv[4] if len(sys.argv) > 4 else "c2"
lib = _lib.load()
if os.environ.get("KMG_SORT_DBG"):                          # 4: generic scatter write-out, 8: register loads instead of the bulk copy
    _lib.check(lib.kmg_tune(b"sort_dbg", int(os.environ["KMG_SORT_DBG"])))
seq = torch.from_numpy(synth.config_c2(L) if wl == "c2" else synth.config_c3(L)).cuda()
for spec in specs:
    parts = [int(x) for x in spec.split(":")]
    rank, shape, rb, bits = parts[:4]
    _lib.check(lib.kmg_tune(b"no_regions", parts[4] if len(parts) > 4 else 0))     # 5th field 1: first pass from a histogram sweep
    _lib.check(lib.kmg_tune(b"sort_cfg", rank))
    _lib.check(lib.kmg_tune(b"sort_shape", shape))
    _lib.check(lib.kmg_tune(b"hash_rb", rb))
    _lib.check(lib.kmg_tune(b"hash_bits", bits))
    for _ in range(2):
        kh.make_kmer_hash(seq, k).free()
    kh.profile(enable=True, reset=True)
    kh.profile(reset=True)
    reps = 5
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        kh.make_kmer_hash(seq, k).free()
    b.record()
    torch.cuda.synchronize()
    prof = kh.profile(enable=False)
    kh.profile(reset=True)
    sp = prof.get("sort_pass", (0, 1, 0))
    line = f"rank {rank} shape {shape} rb {rb} bits {bits:2d} noreg {parts[4] if len(parts) > 4 else 0} build {a.elapsed_time(b) / reps:7.3f} ms | sort_pass {sp[0] / max(sp[1], 1) * 1e3:7.1f} us {sp[2] / max(sp[0], 1e-9) / 1e6:7.0f} GB/s"
    sph = prof.get("sort_pass_hist")
    if sph:
        line += f" | sort_pass_hist {sph[0] / max(sph[1], 1) * 1e3:6.1f}us x{sph[1] // reps} {sph[2] / max(sph[0], 1e-9) / 1e6:5.0f} GB/s"
    for name in ("sort_pass_seq", "hist_all", "hist_seq", "group_detect", "small_fix", "big_fix", "rle", "stats"):
        if name in prof:
            v = prof[name]
            line += f" | {name} {v[0] / max(v[1], 1) * 1e3:6.1f}us"
    print(line, flush=True)
