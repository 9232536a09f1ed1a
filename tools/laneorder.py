"""Run the lane-order self-test (precondition of sort-pass variant RANK 3) and the parity tests under that variant."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmer_hasher_b200 import _lib
L = _lib.load()
f = C.c_uint32(0)
_lib.check(L.kmg_selftest_lane_order(C.byref(f)))
print("lane-order self-test failures:", f.value)
