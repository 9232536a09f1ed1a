"""The R-facing boundary: kmer_hasher_b200/rglue/kmer_hash.c driven through a stand-in R session.
CPU tests cover registration, argument validation and the no-fallback rule; GPU tests cover the
return layouts, lifetime and parity with the oracle through the .Call surface."""
import numpy as np
import pytest

from conftest import random_dna
from rsession import EXTPTRSXP, RError, RSession


@pytest.fixture(scope="module")
def R():
    return RSession()


def test_registered_routines(R):
    # the index entry points, same names and arity as the reference's callMethods table
    with pytest.raises(RError, match="Incorrect number of arguments"):
        R.call("kmer_pair_pos", R.integer(1))
    with pytest.raises(RError, match="not in load table"):
        R.call("count_kmers_fastq", R.integer(1), R.integer(1), R.integer(1))      # read counting: out of scope
    with pytest.raises(RError, match="Incorrect number of arguments"):
        R.call("count_kmers", R.integer(1), R.integer(1))
    with pytest.raises(RError, match="Incorrect number of arguments"):
        R.call("make_kmer_h_index", R.character("ACGT"), R.integer(2))
    with pytest.raises(RError, match="Incorrect number of arguments"):
        R.call("kmer_positions", R.integer(1), R.integer(1), R.integer(1))


def test_argument_errors_match_reference_messages(R):
    with pytest.raises(RError, match="seq_r should be a character vector of length at least one"):
        R.call("make_kmer_h_index", R.integer(1), R.integer(4), R.integer(0))
    with pytest.raises(RError, match="k_r must be an integer vector"):
        R.call("make_kmer_h_index", R.character("ACGTACGT"), R.character("4"), R.integer(0))
    with pytest.raises(RError, match="k must be a positive integer less than 1\\+MAX_K"):
        R.make_kmer_hash("ACGT" * 20, 33)
    with pytest.raises(RError, match="k must be a positive integer less than 1\\+MAX_K"):
        R.make_kmer_hash("ACGT" * 20, 0)
    with pytest.raises(RError, match="the length of the sequence must be at least k"):
        R.make_kmer_hash("ACGT", 4)
    with pytest.raises(RError, match="ptr_r should be an external pointer"):
        R.call("kmer_positions", R.integer(3), R.integer(15))
    with pytest.raises(RError, match="ptr_r should be an external pointer"):
        R.call("sequence_kmer_positions", R.character("x"), R.character("ACGT"), R.integer(2))
    assert R.stub.rstub_protect_depth() == 0 and R.stub.rstub_transient_bytes() == 0


def test_no_gpu_means_a_loud_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    R = RSession()
    with pytest.raises(RError, match="make.kmer.hash failed: .*(no CUDA device|CUDA)"):
        R.make_kmer_hash("ACGTACGTACGTACGT", 4)


@pytest.mark.gpu
def test_call_surface_matches_oracle(R, oracle, test_fa):
    k = 16
    ptr = R.make_kmer_hash(test_fa, k, do_sort=True)
    assert R.typeof(ptr) == EXTPTRSXP
    assert R.strings(R.stub.R_ExternalPtrTag(ptr)) == ["kmer_hash_250930"]
    o = oracle.build(test_fa, k)
    want = o.extract(15)
    got = R.kmer_pos(ptr, 15)
    assert list(got) == ["kmer", "pos", "pair.pos", "count"]
    assert got["kmer"] == [bytes(want["kmer"][i * (k + 1):i * (k + 1) + k]).decode() for i in range(o.U)]
    assert got["pos"].shape == (o.N, 2) and np.array_equal(got["pos"].ravel(), want["pos"])
    assert got["pair.pos"].shape == (o.P, 3) and np.array_equal(got["pair.pos"].ravel(), want["pair_pos"])
    assert np.array_equal(got["count"], want["count"])
    only = R.kmer_pos(ptr, 8)
    assert only["kmer"] is None and only["pos"] is None and only["pair.pos"] is None and len(only["count"]) == o.U
    rows = R.seq_kmer_pos(ptr, test_fa[:20000], k)
    assert np.array_equal(rows.ravel(), o.query(test_fa[:20000], k))
    # kmer.pairs through its .Call entry (kmer_pair_pos): shared k-mers of two indexes
    import oracle as oracle_mod
    other = R.make_kmer_hash(test_fa[10000:30000], k)
    o2 = oracle.build(test_fa[10000:30000], k)
    pr = R.kmer_pairs(ptr, other)
    assert pr.shape[1] == 2 and np.array_equal(pr.ravel(), oracle_mod.pairs_join(want, o2.extract(2 | 8)))
    with pytest.raises(RError, match="ptr_r should be an external pointer"):
        R.call("kmer_pair_pos", ptr, R.integer(1))
    R.stub.rstub_finalize(other)
    assert R.stub.rstub_protect_depth() == 0 and R.stub.rstub_transient_bytes() == 0
    # lifetime: GC finalises once, tolerates a second run, later use is a clean error
    R.stub.rstub_finalize(ptr)
    assert R.stub.R_ExternalPtrAddr(ptr) is None
    R.stub.rstub_finalize(ptr)
    with pytest.raises(RError, match="external pointer is NULL"):
        R.kmer_pos(ptr, 8)


@pytest.mark.gpu
def test_glue_guards_on_device(R, oracle, monkeypatch):
    s = random_dna(5000, 3, p_n=0.002, p_lower=0.2)
    ptr = R.make_kmer_hash(s, 32)
    with pytest.raises(RError, match="should not be longer than 31"):
        R.seq_kmer_pos(ptr, s, 32)                       # the reference's R-level limit
    monkeypatch.setenv("KMERGPU_ALLOW_K32", "1")
    o = oracle.build(s, 32)
    assert np.array_equal(R.seq_kmer_pos(ptr, s[:3000], 32).ravel(), o.query(s[:3000], 32))
    with pytest.raises(RError, match="single sequence"):
        R.call("sequence_kmer_positions", ptr, R.character("ACGT", "ACGT"), R.integer(2))
    with pytest.raises(RError, match="opt_flag_r should be an integer vector of length 1"):
        R.call("kmer_positions", ptr, R.integer(1, 2))
    # a wrong tag is refused like extract_khash_ptr does
    other = R.make_kmer_hash(s, 8)
    tag = R.stub.R_ExternalPtrTag(other)
    import ctypes as C
    C.memmove(R.char_addr(R.stub.STRING_ELT(tag, 0)), b"suffix_hash_2509", 16)
    with pytest.raises(RError, match="External pointer has incorrect tag"):
        R.kmer_pos(other, 8)
    # pair.pos beyond an R matrix: refused before allocating anything
    big = np.full(70000, ord("A"), np.uint8)             # one 12-mer, n = 69989 -> P = 2.4e9 > 2^31-1
    pb = R.make_kmer_hash(big, 12)
    flag = R.integer(4)
    live = R.stub.rstub_live_objects()
    with pytest.raises(RError, match="more than an R matrix can hold"):
        R.call("kmer_positions", pb, flag)
    assert R.stub.rstub_live_objects() == live           # nothing was allocated before the refusal
    assert R.kmer_pos(pb, 8)["count"].tolist() == [69989]
    R.stub.rstub_finalize(ptr); R.stub.rstub_finalize(pb)


@pytest.mark.gpu
def test_default_call_uses_the_grouped_build_and_do_sort_orders_by_key(R, oracle, test_fa):
    """make.kmer.hash(seq, 32) as R calls it (do.sort = FALSE) builds grouped: the k-mers come in another order
    than with do.sort = TRUE, and k-mer for k-mer the strings, counts and position lists are the reference's."""
    k = 32
    o = oracle.build(test_fa, k)
    want = o.extract(1 | 2 | 8)
    want_kmers = [bytes(want["kmer"][i * (k + 1):i * (k + 1) + k]).decode() for i in range(o.U)]
    wpos = want["pos"].reshape(-1, 2)
    starts = np.concatenate([[0], np.cumsum(want["count"])])
    want_lists = {want_kmers[u]: wpos[starts[u]:starts[u + 1], 1].tolist() for u in range(o.U)}
    srt = R.make_kmer_hash(test_fa, k, do_sort=True)
    assert R.kmer_pos(srt, 1)["kmer"] == want_kmers                      # ascending key
    grp = R.make_kmer_hash(test_fa, k)
    got = R.kmer_pos(grp, 1 | 2 | 8)
    assert got["kmer"] != want_kmers and sorted(got["kmer"]) == sorted(want_kmers)
    gstarts = np.concatenate([[0], np.cumsum(got["count"])])
    assert np.array_equal(got["pos"][:, 0], np.repeat(np.arange(1, o.U + 1), got["count"]))
    for u, km in enumerate(got["kmer"]):
        assert got["pos"][gstarts[u]:gstarts[u + 1], 1].tolist() == want_lists[km]
    # probes do not depend on the order of the k-mers
    import os
    os.environ["KMERGPU_ALLOW_K32"] = "1"
    try:
        q = test_fa[3000:9000]
        assert np.array_equal(R.seq_kmer_pos(grp, q, k).ravel(), o.query(q, k))
    finally:
        del os.environ["KMERGPU_ALLOW_K32"]
    R.stub.rstub_finalize(srt); R.stub.rstub_finalize(grp)
    assert R.stub.rstub_protect_depth() == 0
