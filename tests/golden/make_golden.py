"""Regenerates tests/golden/ from the UNMODIFIED reference engine (oracle/_ref/libkmer_ref.so,
i.e. /root/reference/src/kmer_pos.c + kmer_util.c).  Run in the build container only:

    PYTHONPATH=. python tests/golden/make_golden.py

Outputs
  test_fa.seq   the bases of /root/reference/test.fa (BASELINE config 1 input; a data fixture)
  golden.json   * per k in {10,12,16,21,31,32} on test.fa: U, N, P, multi, max count and its key, self-query
                  rows, and SHA-256 of every canonical output array (the arrays themselves would be
                  ~170 MB for pair.pos);
                * full canonical outputs of the small adversarial strings of SURVEY.md Appendix B plus a
                  few more edge cases (N runs, lower case, IUPAC bytes, k = 1 and k = 32, end-of-string rule).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import Reference  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ref = Reference()
    fa = "".join(l.strip() for l in open("/root/reference/test.fa") if not l.startswith(">"))
    with open(os.path.join(HERE, "test_fa.seq"), "w") as fh:
        fh.write(fa + "\n")
    out = {"test_fa": {}, "small": []}
    for k in (10, 12, 16, 21, 31, 32):
        ri = ref.build(fa, k)
        e = ri.extract(15)
        q = ri.query(fa, k)
        cnt = e["count"]
        out["test_fa"][str(k)] = dict(
            U=ri.U, N=ri.N, P=ri.P, multi=int((cnt > 1).sum()), max_count=int(cnt.max()),
            max_key=int(e["keys"][cnt.argmax()]), self_query_rows=len(q) // 2, khash_buckets=ri.buckets,
            sha_keys=sha(e["keys"]), sha_kmer=sha(e["kmer"]), sha_count=sha(cnt), sha_pos=sha(e["pos"]),
            sha_pair_pos=sha(e["pair_pos"]), sha_self_query=sha(q))
        ri.close()
    small = [(4, "ACGTA"), (4, "ACGTNACGT"), (4, "ACGTNACGTA"), (4, "ACGTNACNACGTA"), (4, "NNACGTACGTNN"),
             (3, "acgtRYacgt"), (32, "A" * 40 + "C"), (1, "ACGTNacgtn"), (2, "AAAAAAAA"), (5, "ACGTACGTACGTACGTACGT"),
             (4, "ACGT"), (4, "ACG"), (4, "NNNN"), (4, "ACGTN"), (4, "NACGT"), (4, "nACGTA"), (3, "ACGNNNACGNACG"),
             (32, "ACGT" * 8 + "N" + "ACGT" * 8), (32, "ACGT" * 8 + "N" + "ACGT" * 8 + "A"),
             (31, "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT"), (16, "CCCTAA" * 20), (8, "ACGTacgtNNNNnnnnACGTACGTAC-GT*ACGTAAAC")]
    for k, s in small:
        ri = ref.build(s, k, guard=False)
        e = ri.extract(15)
        wk, wp = ref.windows(s, k)
        out["small"].append(dict(
            k=k, seq=s, U=ri.U, N=ri.N, P=ri.P, keys=[int(x) for x in e["keys"]],
            kmer=[bytes(e["kmer"][i * (k + 1):i * (k + 1) + k]).decode() for i in range(ri.U)],
            count=e["count"].tolist(), pos=e["pos"].tolist(), pair_pos=e["pair_pos"].tolist(),
            self_query=ri.query(s, k).tolist(), window_keys=[int(x) for x in wk], window_pos=wp.tolist()))
        ri.close()
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", len(out["small"]), "small cases and", len(out["test_fa"]), "test.fa cases")


if __name__ == "__main__":
    main()
