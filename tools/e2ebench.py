"""Tuning run: where the end-to-end (host buffers) time of make.kmer.hash + kmer.pos(2|8) goes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import kmer_hasher_b200 as kh
from kmer_hasher_b200 import synth

L, k = 40_000_000, 32
seq_pin = kh.pinned_empty(L, np.uint8); synth.config_c2(L, out=seq_pin)
seq_page = np.array(seq_pin)
ix = kh.make_kmer_hash(seq_pin, k); U, N, P = ix.sizes; ix.free()
pos_pin, cnt_pin = kh.pinned_empty((N, 2), np.int32), kh.pinned_empty(U, np.int32)
pos_page, cnt_page = np.empty((N, 2), np.int32), np.empty(U, np.int32)
def t(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
h = kh.make_kmer_hash(seq_pin, k)
print("build  pinned   %.2f ms" % t(lambda: kh.make_kmer_hash(seq_pin, k).free()))
print("build  pageable %.2f ms" % t(lambda: kh.make_kmer_hash(seq_page, k).free()))
print("pos    pinned   %.2f ms" % t(lambda: kh.kmer_pos(h, 2, out={"pos": pos_pin})))
print("pos    pageable %.2f ms" % t(lambda: kh.kmer_pos(h, 2, out={"pos": pos_page})))
print("count  pinned   %.2f ms" % t(lambda: kh.kmer_pos(h, 8, out={"count": cnt_pin})))
print("count  pageable %.2f ms" % t(lambda: kh.kmer_pos(h, 8, out={"count": cnt_page})))
d = torch.empty(N * 2, dtype=torch.int32, device="cuda")
pt = torch.from_numpy(pos_pin.reshape(-1))
print("torch D2H pinned 320MB %.2f ms" % t(lambda: pt.copy_(d, non_blocking=True)))
pg = torch.from_numpy(pos_page.reshape(-1))
print("torch D2H pageable 320MB %.2f ms" % t(lambda: pg.copy_(d)))
from kmer_hasher_b200 import _lib
_lib.check(_lib.load().kmg_set_stream(torch.cuda.current_stream().cuda_stream))
print("-- on torch's current stream (legacy default stream)")
print("build  pinned   %.2f ms" % t(lambda: kh.make_kmer_hash(seq_pin, k).free()))
print("pos    pinned   %.2f ms" % t(lambda: kh.kmer_pos(h, 2, out={"pos": pos_pin})))
print("count  pinned   %.2f ms" % t(lambda: kh.kmer_pos(h, 8, out={"count": cnt_pin})))
def step():
    hh = kh.make_kmer_hash(seq_pin, k); kh.kmer_pos(hh, 10, out={"pos": pos_pin, "count": cnt_pin}); hh.free()
print("step   pinned   %.2f ms" % t(step, 10))
s2 = torch.cuda.Stream()
with torch.cuda.stream(s2):
    _lib.check(_lib.load().kmg_set_stream(s2.cuda_stream))
    print("-- on a torch side stream")
    print("step   pinned   %.2f ms" % t(step, 10))
