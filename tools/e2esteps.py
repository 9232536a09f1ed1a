import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import kmer_hasher_b200 as kh
from kmer_hasher_b200 import synth, _lib
L, k = 40_000_000, 32
seq_pin = kh.pinned_empty(L, np.uint8); synth.config_c2(L, out=seq_pin)
ix = kh.make_kmer_hash(seq_pin, k); U, N, P = ix.sizes; ix.free()
pos_pin, cnt_pin = kh.pinned_empty((N, 2), np.int32), kh.pinned_empty(U, np.int32)
if len(sys.argv) > 1 and sys.argv[1] == "torchstream":
    _lib.check(_lib.load().kmg_set_stream(torch.cuda.current_stream().cuda_stream))
def step():
    t0 = time.perf_counter()
    hh = kh.make_kmer_hash(seq_pin, k); t1 = time.perf_counter()
    kh.kmer_pos(hh, 10, out={"pos": pos_pin, "count": cnt_pin}); t2 = time.perf_counter()
    hh.free(); torch.cuda.synchronize(); t3 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3
for i in range(12):
    print("step %2d build %.2f extract %.2f free %.2f ms" % ((i,) + step()))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ClockSampler
print("-- with NVML sampler thread")
with ClockSampler(0) as clk:
    for i in range(6):
        print("step %2d build %.2f extract %.2f free %.2f ms" % ((i,) + step()))
print(clk.summary())
print("-- after the sampler")
for i in range(6):
    print("step %2d build %.2f extract %.2f free %.2f ms" % ((i,) + step()))
kh.profile(enable=True, reset=True)
print("-- profiling on")
for i in range(4):
    print("step %2d build %.2f extract %.2f free %.2f ms" % ((i,) + step()))
kh.profile(enable=False); kh.profile(reset=True)
print("-- profiling off again")
for i in range(4):
    print("step %2d build %.2f extract %.2f free %.2f ms" % ((i,) + step()))
