"""Parity at BASELINE.json's FULL sizes, against digests of the reference engine's own outputs.

tests/golden/fullsize.json (made by tests/golden/make_fullsize.py in the build container, from oracle/_ref = the
unmodified /root/reference/src/kmer_pos.c + kmer_util.c) holds, for each configuration, the sizes U, N, P (and M for
the C4 probe) and an order-sensitive digest -- n, sum v_t, sum v_t * (2t + 1) mod 2^64 -- of every canonical output
stream: keys, count, interleaved (i,pos), interleaved (i,x,y), interleaved query rows (i,j).  Here the CUDA path
produces the same outputs through the C ABI, the digests are recomputed ON THE DEVICE, and the numbers must match:

  c3k32  250 Mbp with N gaps, k=32 (the north-star target; grouped build)  keys / count / (i,pos), and
         C4: the 100 Mbp probe of that index, 460 M (i,j) rows
  c3     the same sequence at k=21 (BASELINE config 3)                      keys / count / (i,pos)
  c2     40 Mbp repeat-rich, k=32 (BASELINE config 2)                       keys / count / (i,pos)
  c5     40 Mbp tandem-repeat-heavy, k=12 (BASELINE config 5)               1.47e9 (i,x,y) triples, streamed in chunks

The inputs are regenerated here by the deterministic C generator and their sha256 is checked against the fixture first.
"""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
MASK = (1 << 64) - 1
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fullsize.json")


@pytest.fixture(scope="module")
def full():
    with open(GOLDEN) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def env():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    import kmer_hasher_b200 as kh
    from kmer_hasher_b200 import _lib, synth
    return torch, kh, _lib, synth


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class DevDigest:
    """The digest of oracle/ref_driver.c (dig_add), accumulated on the device over consecutive pieces of a stream."""

    def __init__(self, torch):
        self.torch, self.n, self.sum, self.wsum = torch, 0, 0, 0

    def add(self, x, chunk=1 << 26):
        torch = self.torch
        v = x.reshape(-1)
        for a in range(0, v.numel(), chunk):
            w = v[a:a + chunk].to(torch.int64)
            t = torch.arange(self.n, self.n + w.numel(), device=w.device, dtype=torch.int64)
            self.sum = (self.sum + int(w.sum().item())) & MASK
            self.wsum = (self.wsum + int((w * (2 * t + 1)).sum().item())) & MASK       # int64 arithmetic wraps mod 2^64
            self.n += w.numel()
        return self

    def add_at(self, v, t):
        """values v (int64) that sit at stream indices t (int64), any order"""
        self.sum = (self.sum + int(v.sum().item())) & MASK
        self.wsum = (self.wsum + int((v * (2 * t + 1)).sum().item())) & MASK
        self.n += v.numel()

    def tuple(self):
        return [self.n, self.sum, self.wsum]


def _index_digests(torch, kh, _lib, ix):
    """Digests of the canonical keys / count / (i,pos) streams of an index (k-mers by ascending key, i renumbered),
    computed on the device without re-ordering the rows: row r of k-mer u (index order) is canonical row
    cstart[rank[u]] + (r - ustart[u])."""
    L = _lib.load()
    U, N, P = ix.sizes
    keys = torch.empty(U, dtype=torch.int64, device="cuda")
    _lib.check(L.kmg_kmers_u64(ix._handle(), keys.data_ptr()))
    cnt = torch.empty(U, dtype=torch.int32, device="cuda")
    pos = torch.empty((N, 2), dtype=torch.int32, device="cuda")
    kh.kmer_pos(ix, 2 | 8, out={"pos": pos, "count": cnt})
    flip = torch.tensor(-2**63, dtype=torch.int64, device="cuda")
    skeys, order = torch.sort(keys ^ flip)                     # ascending as unsigned
    skeys ^= flip
    assert bool((skeys[1:] != skeys[:-1]).all())               # distinct
    rank = torch.empty_like(order)
    rank[order] = torch.arange(U, device="cuda", dtype=torch.int64)
    d_keys = DevDigest(torch).add(skeys)
    ccnt = cnt[order].to(torch.int64)
    d_cnt = DevDigest(torch).add(ccnt)
    cstart = torch.cumsum(ccnt, 0) - ccnt                       # first canonical row of each canonical k-mer
    ustart = torch.cumsum(cnt.to(torch.int64), 0) - cnt.to(torch.int64)
    d_pos = DevDigest(torch)
    step = 1 << 26
    for a in range(0, N, step):
        rows = pos[a:a + step]
        u = rows[:, 0].to(torch.int64) - 1
        r = torch.arange(a, a + rows.shape[0], device="cuda", dtype=torch.int64)
        t = cstart[rank[u]] + (r - ustart[u])
        d_pos.add_at(rank[u] + 1, 2 * t)
        d_pos.add_at(rows[:, 1].to(torch.int64), 2 * t + 1)
    # the rows of a k-mer are contiguous and its positions ascend (what makes the formula above the canonical stream)
    same = pos[1:, 0] == pos[:-1, 0]
    assert bool((pos[1:, 1][same] > pos[:-1, 1][same]).all()) and bool((pos[1:, 0] >= pos[:-1, 0]).all())
    return d_keys.tuple(), d_cnt.tuple(), d_pos.tuple()


def _check_index(env, want, seq_dev, k):
    torch, kh, _lib, synth = env
    ix = kh.make_kmer_hash(seq_dev, k)
    assert list(ix.sizes) == [want["U"], want["N"], want["P"]]
    dk, dc, dp = _index_digests(torch, kh, _lib, ix)
    assert dk == want["keys"], "distinct k-mers differ from the reference's"
    assert dc == want["count"], "counts differ from the reference's"
    assert dp == want["pos"], "(i,pos) rows differ from the reference's"
    return ix


def test_c2_40mbp_k32_full_size(env, full):
    torch, kh, _lib, synth = env
    seq = synth.config_c2()
    assert _sha(seq) == full["c2"]["seq_sha256"]
    _check_index(env, full["c2"], torch.from_numpy(seq).cuda(), 32).free()


def test_c3_250mbp_k32_index_and_c4_probe_full_size(env, full):
    torch, kh, _lib, synth = env
    L = _lib.load()
    want = full["c3k32"]
    seq = synth.config_c3()
    assert _sha(seq) == want["seq_sha256"]
    ix = _check_index(env, want, torch.from_numpy(seq).cuda(), 32)
    assert L.kmg_index_order(ix._handle()) == 0                   # the grouped build, as make.kmer.hash runs it
    # C4: seq.kmer.pos of the 100 Mbp query; rows are compared in the emitted order, in chunks (kmg_query_emit_chunk)
    c4 = want["c4"]
    q = synth.config_c4_query(seq, c4["query_bases"])
    assert _sha(q) == c4["query_sha256"]
    dq = torch.from_numpy(q).cuda()
    st, M = C.c_void_p(), C.c_uint64()
    _lib.check(L.kmg_query_begin(ix._handle(), dq.data_ptr(), dq.numel(), 32, C.byref(st), C.byref(M)))
    try:
        assert M.value == c4["M"]
        d = DevDigest(torch)
        step = 1 << 27
        buf = torch.empty((step, 2), dtype=torch.int32, device="cuda")
        for a in range(0, M.value, step):
            n = min(step, M.value - a)
            _lib.check(L.kmg_query_emit_chunk(st, a, n, buf.data_ptr()))
            d.add(buf[:n])
        assert d.tuple() == c4["rows"], "(i,j) rows differ from the reference's seq_kmer_positions"
    finally:
        L.kmg_query_free(st)
    ix.free()


def test_c3_250mbp_k21_full_size(env, full):
    torch, kh, _lib, synth = env
    seq = synth.config_c3()
    assert _sha(seq) == full["c3"]["seq_sha256"]
    _check_index(env, full["c3"], torch.from_numpy(seq).cuda(), 21).free()


def test_c5_pair_triples_full_size(env, full):
    torch, kh, _lib, synth = env
    L = _lib.load()
    want = full["c5"]
    s5 = synth.config_c5()
    assert _sha(s5) == want["seq_sha256"]
    ix = _check_index(env, want, torch.from_numpy(s5).cuda(), 12)
    assert L.kmg_index_order(ix._handle()) == 1                   # k=12: ascending keys, so i needs no renumbering
    P = want["P"]
    assert 10**9 < P < 2**31 - 1
    d = DevDigest(torch)
    step = 1 << 27
    buf = torch.empty((step, 3), dtype=torch.int32, device="cuda")
    for a in range(0, P, step):
        n = min(step, P - a)
        _lib.check(L.kmg_pairs_chunk(ix._handle(), a, n, buf.data_ptr()))
        d.add(buf[:n])
    assert d.tuple() == want["pair_pos"], "(i,x,y) triples differ from the reference's"
    ix.free()
