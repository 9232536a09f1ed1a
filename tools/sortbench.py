"""Tuning run: per-kernel CUDA-event times of make.kmer.hash for each sort-pass configuration.
usage: python tools/sortbench.py [L] [k] [cfg,cfg,...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import kmer_hasher_b200 as kh
from kmer_hasher_b200 import _lib, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfgs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else list(range(12))
dbgs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
if os.environ.get("KMG_HASH_BITS"):
    _lib.check(_lib.load().kmg_tune(b"hash_bits", int(os.environ["KMG_HASH_BITS"])))
lib = _lib.load()
seq = torch.from_numpy(synth.config_c2(L)).cuda()
for cfg, dbg in [(c, d) for c in cfgs for d in dbgs]:
    _lib.check(lib.kmg_tune(b"sort_cfg", cfg))
    _lib.check(lib.kmg_tune(b"sort_dbg", dbg))
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
    line = f"cfg {cfg:2d} dbg {dbg} build {a.elapsed_time(b) / reps:7.3f} ms | sort_pass {sp[0] / max(sp[1], 1) * 1e3:7.1f} us {sp[2] / max(sp[0], 1e-9) / 1e6:7.0f} GB/s"
    sph = prof.get("sort_pass_hist")
    if sph:
        line += f" | sort_pass_hist {sph[0] / max(sph[1], 1) * 1e3:6.1f}us x{sph[1] // reps}"
    for name in ("sort_pass_seq", "hist_all", "hist_seq", "group_detect", "small_fix", "big_fix", "rle", "stats"):
        if name in prof:
            v = prof[name]
            line += f" | {name} {v[0] / max(v[1], 1) * 1e3:6.1f}us"
    print(line, flush=True)
