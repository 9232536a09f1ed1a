"""CPU tests: the plain-C oracle (oracle/kmer_oracle.c) is pinned against
  (1) golden vectors generated from the unmodified reference engine (tests/golden/), and
  (2) the reference engine itself (oracle/_ref) on seeded random inputs, when it is available.
"""
import numpy as np
import pytest

from conftest import random_dna, sha

FIELDS = ("keys", "kmer", "count", "pos", "pair_pos")


def test_golden_test_fa(oracle, golden, test_fa):
    assert len(test_fa) == 59940
    for k, g in golden["test_fa"].items():
        ix = oracle.build(test_fa, int(k))
        assert (ix.U, ix.N, ix.P) == (g["U"], g["N"], g["P"])
        e = ix.extract(15)
        assert int(e["count"].max()) == g["max_count"]
        assert int(e["keys"][e["count"].argmax()]) == g["max_key"]
        assert int((e["count"] > 1).sum()) == g["multi"]
        for f in FIELDS:
            assert sha(e[f]) == g["sha_" + f], (k, f)
        q = ix.query(test_fa, int(k))
        assert len(q) // 2 == g["self_query_rows"]
        assert sha(q) == g["sha_self_query"]


def test_golden_small(oracle, golden):
    for c in golden["small"]:
        k, s = c["k"], c["seq"]
        wk, wp = oracle.windows(s, k)
        assert wk.tolist() == c["window_keys"] and wp.tolist() == c["window_pos"], (k, s)
        ix = oracle.build(s, k, guard=False)
        assert (ix.U, ix.N, ix.P) == (c["U"], c["N"], c["P"])
        e = ix.extract(15)
        assert e["keys"].tolist() == c["keys"]
        assert e["count"].tolist() == c["count"]
        assert e["pos"].tolist() == c["pos"]
        assert e["pair_pos"].tolist() == c["pair_pos"]
        kmers = [bytes(e["kmer"][i * (k + 1):i * (k + 1) + k]).decode() for i in range(ix.U)]
        assert kmers == c["kmer"]
        assert ix.query(s, k).tolist() == c["self_query"]


def test_guards(oracle):
    with pytest.raises(ValueError, match="less than 1\\+MAX_K"):
        oracle.build("ACGT" * 20, 33)
    with pytest.raises(ValueError, match="less than 1\\+MAX_K"):
        oracle.build("ACGT" * 20, 0)
    with pytest.raises(ValueError, match="at least k"):
        oracle.build("ACGT", 4)


@pytest.mark.parametrize("k", [1, 3, 8, 12, 16, 21, 31, 32])
def test_against_reference_engine(oracle, reference, k):
    for seed, kw in enumerate([dict(), dict(p_n=0.01), dict(p_n=0.002, p_lower=0.3, p_other=0.01, n_runs=5),
                               dict(n_runs=30)]):
        s = random_dna(20000 if k > 3 else 3000, 100 + seed, **kw)
        a = reference.build(s, k)
        b = oracle.build(s, k)
        assert (a.U, a.N, a.P) == (b.U, b.N, b.P)
        ea, eb = a.extract(15), b.extract(15)
        for f in FIELDS:
            assert np.array_equal(ea[f], eb[f]), (k, seed, f)
        q = random_dna(5000, 900 + seed, **kw)
        q[1000:3000] = s[500:2500]
        assert np.array_equal(a.query(q, k), b.query(q, k))
        wa, wb = reference.windows(s, k), oracle.windows(s, k)
        assert np.array_equal(wa[0], wb[0]) and np.array_equal(wa[1], wb[1])


def test_raw_bucket_order_is_a_permutation(reference):
    """The reference's own (bucket-order) output equals the canonical one up to the k-mer permutation."""
    s = random_dna(5000, 7, p_n=0.003)
    ix = reference.build(s, 6)
    raw, can = ix.extract_raw(15), ix.extract(15)
    order = np.argsort(raw["keys"], kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    assert np.array_equal(raw["keys"][order], can["keys"])
    assert np.array_equal(raw["count"][order], can["count"])
    pos = raw["pos"].reshape(-1, 2).copy()
    pos[:, 0] = rank[pos[:, 0] - 1] + 1
    pos = pos[np.argsort(pos[:, 0], kind="stable")]
    assert np.array_equal(pos.ravel(), can["pos"])
    pp = raw["pair_pos"].reshape(-1, 3).copy()
    pp[:, 0] = rank[pp[:, 0] - 1] + 1
    pp = pp[np.argsort(pp[:, 0], kind="stable")]
    assert np.array_equal(pp.ravel(), can["pair_pos"])


def test_query_k_independent_of_index_k(oracle, reference):
    s = random_dna(4000, 11)
    a, b = reference.build(s, 8), oracle.build(s, 8)
    for kq in (4, 8, 12):
        assert np.array_equal(a.query(s[:2000], kq), b.query(s[:2000], kq))


def test_pairs_join_restatement_matches_reference_tables(oracle, reference):
    """kmer.pairs: the numpy restatement (oracle.pairs_join) against ref_pairs_join, which walks the
    reference's own khash tables (kmer_pair_pos with its missing kh_exist added)."""
    import oracle as oracle_mod
    from kmer_hasher_b200 import synth
    a_seq = synth.config_c3(60_000)
    b_seq = synth.config_c4_query(a_seq, 30_000)
    for ka, kb in ((12, 12), (16, 16), (9, 7)):
        ra, rb = reference.build(a_seq, ka), reference.build(b_seq, kb)
        want = ra.pairs_join(rb)
        got = oracle_mod.pairs_join(ra.extract(2 | 8), rb.extract(2 | 8))
        assert np.array_equal(got, want)
        if ka == kb:
            assert len(want) > 0
        # symmetry: swapping the indexes swaps the columns (as a multiset of rows)
        back = rb.pairs_join(ra).reshape(-1, 2)[:, ::-1]
        w2 = want.reshape(-1, 2)
        assert np.array_equal(back[np.lexsort((back[:, 1], back[:, 0]))], w2[np.lexsort((w2[:, 1], w2[:, 0]))])
        ra.close(); rb.close()
    # known answer: a = ACGTACGT (k=4): ACGT@{1,5}, CGTA@2, GTAC@3, TACG@4 ; b = TTACGTT: TTAC@1, TACG@2, ACGT@3, CGTT@4
    ra, rb = reference.build("ACGTACGTA"[:8] + "A", 4), reference.build("TTACGTT", 4)
    rows = ra.pairs_join(rb).reshape(-1, 2)
    # shared: ACGT (a@1,5 ; b@3), TACG (a@4 ; b@2); keys ascending: ACGT = 0b00011011 = 27... order by key
    assert sorted(map(tuple, rows.tolist())) == [(1, 3), (4, 2), (5, 3)]
    ra.close(); rb.close()


def test_digest_implementations_agree(reference):
    """The three statements of the full-size digest -- C over the reference's tables (oracle/ref_driver.c), numpy
    (oracle.digest) and torch (tests/test_fullsize_gpu.DevDigest, run on the device by the GPU test) -- give the same
    numbers, including the piecewise and scattered accumulation the GPU test uses."""
    import torch
    from kmer_hasher_b200 import synth
    from oracle import digest
    from test_fullsize_gpu import DevDigest
    seq = synth.config_c3(150_000, tail_k=12)
    r = reference.build(seq, 12)
    e, d = r.extract(15), r.digest(15)
    for name in ("keys", "count", "pos", "pair_pos"):
        assert d[name] == digest(e[name])
        x = torch.from_numpy(e[name].view(np.int64) if e[name].dtype == np.uint64 else e[name])
        assert tuple(DevDigest(torch).add(x[:1000]).add(x[1000:], chunk=4096).tuple()) == d[name]
    pos = torch.from_numpy(e["pos"]).to(torch.int64)
    perm = torch.randperm(pos.numel())
    dd = DevDigest(torch)
    dd.add_at(pos[perm], perm)
    assert tuple(dd.tuple()) == d["pos"]
    keys, cnt = e["keys"], e["count"].astype(np.uint64)
    rows = e["pos"].reshape(-1, 2)
    with np.errstate(over="ignore"):
        assert d["bind"] == (int((keys * cnt).sum(dtype=np.uint64)),
                             int((keys[rows[:, 0] - 1] * rows[:, 1].astype(np.uint64)).sum(dtype=np.uint64)))
    q = synth.config_c4_query(seq, 40_000)
    n, dq = r.query_digest(q, 12)
    assert dq == digest(r.query(q, 12)) and n == dq[0] // 2
    r.close()


def test_restated_count_kmers_is_pinned_to_seq_to_hash(reference):
    """seq_to_counts lives in the R glue file (needs R.h) and is restated in oracle/ref_driver.c.  Pin: with one
    source, the count of every k-mer equals the length of its position list in the index the UNMODIFIED seq_to_hash
    builds from the same sequence (same window rule, same keys); with several sources and calls the columns add up."""
    from kmer_hasher_b200 import synth
    a = synth.config_c3(80_000, tail_k=9)
    b = synth.config_c2(60_000)
    for k in (4, 9, 21, 32):
        ix = reference.build(a, k)
        lists = ix.extract(8)
        ct = reference.count_kmers(a, k, 0, 1)
        e = ct.extract(2 | 8)
        assert np.array_equal(e["keys"], lists["keys"])
        assert np.array_equal(e["pos"].reshape(-1, 2)[:, 1], lists["count"])      # (i, count) rows
        assert (e["count"] == 1).all() and ct.new_kmers == ix.U
        ct.close()
        # two sources, three calls: column sums equal per-sequence list lengths
        ct = reference.count_kmers([a, "ACGT"], k, 0, 2)                           # "ACGT" (length <= k for k >= 4) is skipped
        ct = reference.count_kmers(b, k, 1, 2, ct)
        ct = reference.count_kmers(b[:20_000], k, 0, 2, ct)
        e = ct.extract(2)
        m = dict(zip(e["keys"].tolist(), e["pos"].reshape(-1, 2, 2)[:, :, 1].tolist()))
        ib, ib2 = reference.build(b, k), reference.build(b[:20_000], k)
        la, lb, lb2 = (dict(zip(x["keys"].tolist(), x["count"].tolist())) for x in (lists, ib.extract(8), ib2.extract(8)))
        assert set(m) == set(la) | set(lb)
        for key, (c0, c1) in m.items():
            assert c0 == la.get(key, 0) + lb2.get(key, 0) and c1 == lb.get(key, 0)
        for x in (ix, ib, ib2, ct):
            x.close()
    with pytest.raises(ValueError):
        reference.count_kmers(a, 21, 2, 2)
