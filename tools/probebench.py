"""Tuning run: seq.kmer.pos on BASELINE config 4 (100 Mbp query vs the 250 Mbp index, k=32): first probe (key-table build)
and steady state, per-kernel CUDA-event times.  usage: python tools/probebench.py [L] [Lq] [k]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import kmer_hasher_b200 as kh
from kmer_hasher_b200 import _lib, synth

L = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000_000
Lq = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 32
lib = _lib.load()
seq = synth.config_c3(L)
q = synth.config_c4_query(seq, Lq)
dseq, dq = torch.from_numpy(seq).cuda(), torch.from_numpy(q).cuda()


def show(tag):
    for n, v in sorted(kh.profile(reset=True).items()):
        if v[0] > 0:
            print(f"    [{tag}] {n:18s} {v[0]:8.3f} ms {v[1]:3d} launches {v[2] / v[0] / 1e6 if v[2] else 0:7.0f} GB/s")


def timed(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = fn(); b.record(); torch.cuda.synchronize()
    return r, a.elapsed_time(b)


for cas in (0, 1):
    _lib.check(lib.kmg_tune(b"hash_cas", cas))
    for rep in range(2):
        ix = kh.make_kmer_hash(dseq, k)
        kh.profile(enable=True, reset=True)
        kh.profile(reset=True)
        st, M = C.c_void_p(), C.c_uint64()
        _, ms = timed(lambda: _lib.check(lib.kmg_query_begin(ix._handle(), dq.data_ptr(), Lq, k, C.byref(st), C.byref(M))))
        print(f"hash_cas={cas} rep {rep}: first probe {ms:.2f} ms (M={M.value})")
        show("first")
        lib.kmg_query_free(st)
        if rep == 1:
            rows = torch.empty((M.value, 2), dtype=torch.int32, device="cuda")
            for r2 in range(3):
                _, mb = timed(lambda: _lib.check(lib.kmg_query_begin(ix._handle(), dq.data_ptr(), Lq, k, C.byref(st), C.byref(M))))
                _, me = timed(lambda: _lib.check(lib.kmg_query_emit(st, rows.data_ptr())))
                lib.kmg_query_free(st)
            print(f"  steady: begin {mb:.2f} ms emit {me:.2f} ms -> {(Lq - k + 1) / (mb + me) / 1e6:.2f} G k-mers/s queried")
            show("steady x3")
            del rows
        kh.profile(enable=False)
        ix.free()
