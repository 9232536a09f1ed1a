import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch, kmer_hasher_b200 as kh
from kmer_hasher_b200 import synth
kh.profile(enable=True, reset=True)
for L in (200_000_000, 600_000_000, 1_000_000_000, 1_100_000_000, 1_500_000_000):
    for kind in ("plain", "gaps"):
        s = synth.generate(L, 0xB16, lower=0.2) if kind == "plain" else synth.generate(L, 0xB16, lower=0.2, n_gaps=200, gap_max=200000, n_single=5000)
        d = torch.from_numpy(s).cuda()
        for rep in range(2):
            kh.profile(reset=True)
            ix = kh.make_kmer_hash(d, 32); ix.free()
            p = kh.profile(reset=True)
        print(L, kind, "hist_all %.3f ms" % p["hist_all"][0], "sort_pass_seq %.3f" % p["sort_pass_seq"][0], flush=True)
        del d
