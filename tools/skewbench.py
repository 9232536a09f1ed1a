import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np, torch, kmer_hasher_b200 as kh
from kmer_hasher_b200 import synth
d5 = torch.from_numpy(synth.config_c5()).cuda()
d2 = torch.from_numpy(synth.config_c2()).cuda()
kh.profile(enable=True, reset=True)
for name, d, k in (("c2 k32", d2, 32), ("c5 k12", d5, 12), ("c5 k12", d5, 12), ("c5 k12", d5, 12), ("c2 k12", d2, 12), ("c2 k12", d2, 12), ("c5 k32", d5, 32), ("c5 k32", d5, 32)):
    torch.cuda.synchronize(); t0=time.perf_counter(); ix = kh.make_kmer_hash(d, k); torch.cuda.synchronize(); ms=(time.perf_counter()-t0)*1e3
    p = kh.profile(reset=True); ix.free()
    print(name, "%.2f ms |" % ms, " ".join(f"{n}={v[0]:.3f}" for n, v in sorted(p.items())))
