import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch, kmer_hasher_b200 as kh
from kmer_hasher_b200 import synth
kh.profile(enable=True, reset=True)
for frac in (0.02, 0.025, 0.03, 0.035, 0.04):
    s = synth.generate(40_000_000, 0xC5, tandem=frac, tandem_unit_max=40, tandem_len_max=60_000)
    ix = kh.make_kmer_hash(torch.from_numpy(s).cuda(), 12); print(frac, ix.sizes, ix.sizes[2] < 2**31-1); ix.free()
for n, v in sorted(kh.profile(reset=True).items()): print(n, round(v[0]/5,3), 'ms')
s = synth.config_c3(250_000_000); q = synth.config_c4_query(s, 100_000_000)
import ctypes as C
from kmer_hasher_b200 import _lib
L=_lib.load(); ix = kh.make_kmer_hash(torch.from_numpy(s).cuda(), 32); dq=torch.from_numpy(q).cuda()
st, M = C.c_void_p(), C.c_uint64(); _lib.check(L.kmg_query_begin(ix._handle(), dq.data_ptr(), len(q), 32, C.byref(st), C.byref(M))); print('C4 M', M.value, M.value < 2**31-1, ix.sizes)
