"""Tuning run: per-phase clock64 timeline of the tiles of one sort pass (third pass of a k=32 build)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import kmer_hasher_b200 as kh
from kmer_hasher_b200 import _lib, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
cfg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tile = int(sys.argv[3]) if len(sys.argv) > 3 else 6144
lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
seq = torch.from_numpy(synth.config_c2(L)).cuda()
_lib.check(lib.kmg_tune(b"sort_cfg", 3 if cfg else 0))
for _ in range(2):
    kh.make_kmer_hash(seq, 32).free()
ntr = (L + 2047) // 2048
_lib.check(lib.kmg_tune(b"sort_trace", ntr))
kh.make_kmer_hash(seq, 32).free()
torch.cuda.synchronize()
buf = np.zeros((ntr, 8), np.uint64)
assert raw.kmg_trace_read(buf.ctypes.data_as(C.c_void_p), ntr) == 0
nt = (L - 31 + tile - 1) // tile
t = buf[:nt].astype(np.int64)
ok = (t[:, 0] > 0) & (t[:, 7] > 0)
t = t[ok]
names = ["ticket+hist scan", "first key arrives", "rank (thread 0's warp)", "wait other warps (sync)", "totals+scan+regroup",
         "look-back + sync", "write-out issue"]
d = np.diff(t, axis=1).astype(np.float64)
print(f"tiles traced {len(t)}; cycles per phase: mean / median / p90")
for i, n in enumerate(names):
    print(f"  {n:28s} {d[:, i].mean():9.0f} {np.median(d[:, i]):9.0f} {np.percentile(d[:, i], 90):9.0f}")
tot = (t[:, 7] - t[:, 0]).astype(np.float64)
print(f"  {'total':28s} {tot.mean():9.0f} {np.median(tot):9.0f} {np.percentile(tot, 90):9.0f}")
