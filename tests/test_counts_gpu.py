"""count.kmers and spectra against the oracle: restated seq_to_counts (src/kmer_hash.c:185-252)
on the reference's own khash, read back the way R reads a count table (kmer.pos walks the counters)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kh():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    import kmer_hasher_b200 as kh
    return kh


def _seqs():
    from kmer_hasher_b200 import synth
    a = synth.config_c3(120_000, tail_k=16)
    b = synth.config_c2(90_000)
    b[1000:1400] = np.frombuffer(b"CCCTAA" * 66 + b"CCCT", np.uint8)      # telomere-like repeat shared with c
    c = a.copy()[:70_000]
    c[5000:5400] = b[1000:1400]
    short = np.frombuffer(b"ACGTACGTAC", np.uint8)                          # length <= k: skipped, as in the reference
    return a, b, c, short


@pytest.mark.parametrize("k", [5, 16, 21, 32])
def test_count_kmers_matches_the_reference(kh, reference, k):
    a, b, c, short = _seqs()
    sn = 3
    ptr = kh.count_kmers([a, short], (k, 0, sn))
    ptr = kh.count_kmers(b, (k, 1, sn), ptr)
    ptr = kh.count_kmers([c, a[:50_000]], (k, 2, sn), ptr)
    ptr = kh.count_kmers(b[:30_000], (k, 0, sn), ptr)                        # a second batch into a filled column
    ref = reference.count_kmers([a, short], k, 0, sn)
    ref = reference.count_kmers(b, k, 1, sn, ref)
    ref = reference.count_kmers([c, a[:50_000]], k, 2, sn, ref)
    ref = reference.count_kmers(b[:30_000], k, 0, sn, ref)
    want = ref.extract(15)
    got = kh.kmer_pos(ptr, 15)
    U = ref.U
    assert ptr.sizes[0] == U and ptr.kmer_count == ref.new_kmers
    assert np.array_equal(kh.kmer_keys(ptr), want["keys"])
    assert np.array_equal(got["kmer"].view(np.uint8).reshape(U, k), want["kmer"].reshape(U, k + 1)[:, :k])
    assert np.array_equal(got["pos"].ravel(), want["pos"])                   # rows (i, count of source s)
    assert np.array_equal(got["pair.pos"].ravel(), want["pair_pos"])
    assert np.array_equal(got["count"], want["count"]) and (got["count"] == sn).all()
    # spectra: per source and summed, clamped at max_count
    m = want["pos"].reshape(U, sn, 2)[:, :, 1].astype(np.int64)
    for source, col in ((0, m[:, 0]), (2, m[:, 2]), (None, m.sum(1))):
        for max_count in (3, 50, 100000):
            spec = kh.kmer_spectrum(ptr, max_count, source)
            assert np.array_equal(spec, np.bincount(np.minimum(col, max_count), minlength=max_count + 1).astype(np.float64))
    with pytest.raises(ValueError):
        kh.seq_kmer_pos(ptr, a, k if k < 32 else 31)
    with pytest.raises(ValueError):
        kh.count_kmers(a, (k + 1 if k < 32 else 31, 0, sn), ptr)
    with pytest.raises(ValueError):
        kh.count_kmers(a, (k, 3, sn), ptr)
    ptr.free(); ref.close()


def test_index_spectrum_is_the_histogram_of_list_lengths(kh, oracle):
    a, b, _, _ = _seqs()
    for seq, k in ((a, 21), (b, 12), (b, 32)):
        ix = kh.make_kmer_hash(seq, k)
        cnt = oracle.build(seq, k).extract(8)["count"].astype(np.int64)
        for max_count in (1, 7, 5000):
            assert np.array_equal(kh.kmer_spectrum(ix, max_count),
                                  np.bincount(np.minimum(cnt, max_count), minlength=max_count + 1).astype(np.float64))
        ix.free()


def test_count_kmers_through_the_r_glue(kh, reference):
    """.Call("count_kmers", ptr, params, seq) then kmer.pos(ptr, 1+2+8), as test.R:340-343 does."""
    from rsession import RSession, RError
    a, b, _, short = _seqs()
    R = RSession()
    k, sn = 21, 2
    ptr = R.call("count_kmers", R.nil, R.integer(k, 0, sn), R.character(a, short))
    ptr2 = R.call("count_kmers", ptr, R.integer(k, 1, sn), R.character(b))
    assert ptr2 == ptr
    got = R.kmer_pos(ptr, 1 + 2 + 8)
    ref = reference.count_kmers([a, short], k, 0, sn)
    ref = reference.count_kmers(b, k, 1, sn, ref)
    want = ref.extract(1 | 2 | 8)
    U = ref.U
    assert got["pair.pos"] is None
    assert np.array_equal(got["pos"].ravel(), want["pos"])
    assert np.array_equal(got["count"], want["count"])
    assert got["kmer"] == [bytes(r).decode() for r in want["kmer"].reshape(U, k + 1)[:, :k]]
    with pytest.raises(RError, match="mismatch between specified k"):
        R.call("count_kmers", ptr, R.integer(k - 1, 0, sn), R.character(a))
    with pytest.raises(RError, match="source_n must be larger"):
        R.call("count_kmers", ptr, R.integer(k, 2, sn), R.character(a))
    with pytest.raises(RError, match="k-mer counts"):
        R.call("sequence_kmer_positions", ptr, R.character(a), R.integer(k))
    ix = R.make_kmer_hash(a, k)
    with pytest.raises(RError, match="position index"):
        R.call("count_kmers", ix, R.integer(k, 0, sn), R.character(a))
    R.stub.rstub_finalize(ptr); R.stub.rstub_finalize(ix)
    ref.close()
