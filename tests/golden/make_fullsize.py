#!/usr/bin/env python
"""Generates tests/golden/fullsize.json: digests of the reference engine's outputs at BASELINE.json's FULL sizes.

Run in the build container (needs /root/reference through oracle/_ref, ~40 GB of host memory, ~10 min):

    python tests/golden/make_fullsize.py [c2] [c3k32] [c3] [c5]

For every configuration the UNMODIFIED reference engine (oracle/_ref = /root/reference/src/kmer_pos.c +
kmer_util.c) indexes the synthetic sequence of that configuration (kmer_hasher_b200.synth, deterministic C
generator: same bytes on every host, pinned by the sha256 recorded here) and oracle/ref_driver.c's
ref_digest_canonical / ref_query_digest walk its hash table in canonical order and record, for each output
stream (keys, count, interleaved (i,pos), interleaved (i,x,y), interleaved query rows (i,j)):

    n, sum v_t mod 2^64, sum v_t * (2t + 1) mod 2^64

tests/test_fullsize_gpu.py recomputes the same numbers from the CUDA path's outputs on the device.
The arrays themselves (2 GB of pos rows, 17.6 GB of pair rows) are never stored.
"""
import hashlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from kmer_hasher_b200 import synth  # noqa: E402
from oracle import Reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def index_digests(ref, seq, k, flag):
    t0 = time.time()
    r = ref.build(seq, k)
    d = {"k": k, "U": r.U, "N": r.N, "P": r.P, "reference_build_seconds": round(r.build_seconds, 2)}
    d.update({n: list(v) for n, v in r.digest(flag).items()})
    print(f"  k={k}: U={r.U} N={r.N} P={r.P} build {r.build_seconds:.1f}s, digests after {time.time() - t0:.0f}s", flush=True)
    return r, d


def main():
    what = set(sys.argv[1:]) or {"c2", "c3k32", "c3", "c5"}
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    ref = Reference()
    if "c2" in what:
        print("c2: 40 Mbp repeat-rich, k=32", flush=True)
        seq = synth.config_c2()
        r, d = index_digests(ref, seq, 32, 2 | 8)
        r.close()
        res["c2"] = {"bases": int(seq.size), "seq_sha256": sha(seq), **d}
    if what & {"c3k32", "c3"}:
        seq = synth.config_c3()
        seq_sha = sha(seq)
    if "c3k32" in what:
        print("c3k32: 250 Mbp with N gaps, k=32 (+ C4: 100 Mbp query)", flush=True)
        r, d = index_digests(ref, seq, 32, 2 | 8)
        q = synth.config_c4_query(seq, 100_000_000)
        M, dq = r.query_digest(q, 32)
        print(f"  C4 query: M={M} rows in {r.query_seconds:.1f}s", flush=True)
        r.close()
        res["c3k32"] = {"bases": int(seq.size), "seq_sha256": seq_sha, **d,
                        "c4": {"query_bases": int(q.size), "query_sha256": sha(q), "k": 32, "M": M, "rows": list(dq),
                               "reference_query_seconds": round(r.query_seconds, 2)}}
    if "c3" in what:
        print("c3: 250 Mbp with N gaps, k=21", flush=True)
        r, d = index_digests(ref, seq, 21, 2 | 8)
        r.close()
        res["c3"] = {"bases": int(seq.size), "seq_sha256": seq_sha, **d}
    if "c5" in what:
        print("c5: 40 Mbp tandem-repeat-heavy, k=12, pair.pos", flush=True)
        s5 = synth.config_c5()
        r, d = index_digests(ref, s5, 12, 2 | 4 | 8)
        r.close()
        res["c5"] = {"bases": int(s5.size), "seq_sha256": sha(s5), **d}
    res["_how"] = "python tests/golden/make_fullsize.py (reference engine = oracle/_ref, digests = oracle/ref_driver.c)"
    json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
