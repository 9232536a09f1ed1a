"""GPU parity tests: the CUDA library, called through its C ABI (ctypes mirror of the R API), must be
bit-exact with the oracle on the same inputs -- k-mer set, counts, per-k-mer positions, pair triples,
query (i,j) rows."""
import numpy as np
import pytest

from conftest import random_dna, sha

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kh():
    import kmer_hasher_b200 as kh
    return kh


def kmers_of(e, U, k):
    return e["kmer"].reshape(U, k + 1)[:, :k] if e["kmer"].ndim == 1 and e["kmer"].dtype == np.uint8 else e["kmer"]


def compare_index(kh, oracle, seq, k, flags=15, guard=True):
    o = oracle.build(seq, k, guard=guard)
    g = kh.make_kmer_hash(seq, k)
    assert g.sizes == (o.U, o.N, o.P), (g.sizes, (o.U, o.N, o.P))
    eo = o.extract(flags)
    eg = kh.kmer_pos(g, flags, canonical=True)
    assert np.array_equal(kh.kmer_keys(g, canonical=True), eo["keys"])
    if flags & 1:
        want = np.ascontiguousarray(eo["kmer"].reshape(o.U, k + 1)[:, :k]).view(f"S{k}").ravel()
        assert np.array_equal(eg["kmer"], want)
    if flags & 2:
        assert np.array_equal(eg["pos"].ravel(), eo["pos"])
    if flags & 4:
        assert np.array_equal(eg["pair.pos"].ravel(), eo["pair_pos"])
    if flags & 8:
        assert np.array_equal(eg["count"], eo["count"])
    return g, o


def test_golden_small(kh, oracle, golden):
    for c in golden["small"]:
        k, s = c["k"], c["seq"]
        if len(s) <= k:      # the R-level guard rejects these; the C ABI itself must agree with the C core
            import ctypes as C
            from kmer_hasher_b200 import _lib
            L = _lib.load()
            h = C.c_void_p()
            _lib.check(L.kmg_build(s.encode(), len(s), k, C.byref(h)))
            U, N, P = C.c_uint64(), C.c_uint64(), C.c_uint64()
            _lib.check(L.kmg_sizes(h, C.byref(U), C.byref(N), C.byref(P)))
            assert (U.value, N.value, P.value) == (c["U"], c["N"], c["P"]) == (0, 0, 0)
            L.kmg_free(h)
            continue
        g = kh.make_kmer_hash(s, k)
        assert g.sizes == (c["U"], c["N"], c["P"]), (k, s)
        e = kh.kmer_pos(g, 15, canonical=True)
        assert kh.kmer_keys(g, canonical=True).tolist() == c["keys"]
        assert [x.decode() for x in e["kmer"]] == c["kmer"]
        assert e["count"].tolist() == c["count"]
        assert e["pos"].ravel().tolist() == c["pos"]
        assert e["pair.pos"].ravel().tolist() == c["pair_pos"]
        q = kh.seq_kmer_pos(g, s, k, allow_k32=True)
        assert q.ravel().tolist() == c["self_query"], (k, s)


@pytest.mark.parametrize("k", [10, 12, 16, 21, 31, 32])
def test_golden_test_fa(kh, golden, test_fa, k):
    """BASELINE config 1: test.fa, make.kmer.hash + kmer.pos(opt.flag=15) (+ self seq.kmer.pos)."""
    g = golden["test_fa"][str(k)]
    ix = kh.make_kmer_hash(test_fa, k)
    assert ix.sizes == (g["U"], g["N"], g["P"])
    e = kh.kmer_pos(ix, 15, canonical=True)
    assert sha(kh.kmer_keys(ix, canonical=True)) == g["sha_keys"]
    assert sha(e["count"]) == g["sha_count"]
    assert sha(e["pos"]) == g["sha_pos"]
    assert sha(e["pair.pos"]) == g["sha_pair_pos"]
    U = g["U"]
    buf = np.zeros((U, k + 1), np.uint8)
    buf[:, :k] = e["kmer"].view(np.uint8).reshape(U, k)
    assert sha(buf) == g["sha_kmer"]
    q = kh.seq_kmer_pos(ix, test_fa, k, allow_k32=True)
    assert len(q) == g["self_query_rows"]
    assert sha(q) == g["sha_self_query"]


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 8, 11, 13, 16, 17, 21, 24, 27, 31, 32])
def test_random_with_breakers(kh, oracle, k):
    n = 60000 if k >= 8 else 6000       # small k => huge pair counts
    for seed, kw in enumerate([dict(), dict(p_n=0.01), dict(p_n=0.001, p_lower=0.3, p_other=0.01, n_runs=8),
                               dict(n_runs=100)]):
        s = random_dna(n, 1000 * k + seed, **kw)
        g, o = compare_index(kh, oracle, s, k)
        q = random_dna(20000 if k >= 8 else 2000, 77 + seed, **kw)
        m = len(q) // 3
        q[m:2 * m] = s[100:100 + m]
        assert np.array_equal(kh.seq_kmer_pos(g, q, k, allow_k32=True).ravel(), o.query(q, k))


@pytest.mark.parametrize("n", [33, 64, 100, 8191, 8192, 8193, 8192 + 31, 16384 + 17, 50001])
def test_tile_boundaries(kh, oracle, n):
    """Lengths around the 8192-window tile and 16-byte group sizes, breakers at the edges."""
    k = 32 if n > 40 else 8
    s = random_dna(n, n)
    for variant in range(4):
        t = s.copy()
        if variant == 1:
            t[0] = ord("N"); t[-1] = ord("n")
        elif variant == 2:
            t[n - k - 1] = ord("N")          # final run of length exactly k: dropped (end-of-string rule)
        elif variant == 3:
            t[n - k - 2] = ord("N")          # final run of length k+1: two windows... one kept
            t[8191 % n] = ord("n")
        g, o = compare_index(kh, oracle, t, k, flags=2 | 8)
        assert np.array_equal(kh.seq_kmer_pos(g, t, k, allow_k32=True).ravel(), o.query(t, k))


def test_skew_homopolymer_and_microsatellite(kh, oracle):
    s = np.concatenate([np.full(30000, ord("A"), np.uint8), np.frombuffer(b"CA" * 10000, np.uint8),
                        np.frombuffer(b"CCCTAA" * 4000, np.uint8), random_dna(20000, 5)])
    for k in (4, 12, 32):
        g, o = compare_index(kh, oracle, s, k, flags=1 | 2 | 8)
        assert g.sizes[2] == o.P
    # pair rows on a lighter version (pairs grow quadratically)
    s2 = np.concatenate([np.full(700, ord("A"), np.uint8), np.frombuffer(b"CA" * 600, np.uint8), random_dna(3000, 6)])
    compare_index(kh, oracle, s2, 6, flags=15)


def test_pairs_chunked_equals_whole(kh, oracle, test_fa):
    import ctypes as C
    from kmer_hasher_b200 import _lib
    L = _lib.load()
    ix = kh.make_kmer_hash(test_fa, 16)
    U, N, P = ix.sizes
    whole = kh.kmer_pos(ix, 4)["pair.pos"]              # the index's own order: chunks are slices of it
    rng = np.random.default_rng(3)
    for _ in range(6):
        first = int(rng.integers(0, P))
        n = int(min(P - first, rng.integers(1, 3_000_000)))
        out = np.empty((n, 3), np.int32)
        _lib.check(L.kmg_pairs_chunk(ix._handle(), first, n, out.ctypes.data))
        assert np.array_equal(out, whole[first:first + n])


def test_query_k_differs_from_index_k(kh, oracle):
    s = random_dna(30000, 21, p_n=0.002)
    g = kh.make_kmer_hash(s, 8)
    o = oracle.build(s, 8)
    for kq in (4, 8, 12, 31):
        assert np.array_equal(kh.seq_kmer_pos(g, s[:9000], kq).ravel(), o.query(s[:9000], kq))


def test_r_level_guards(kh):
    s = "ACGT" * 100
    ix = kh.make_kmer_hash([s, "ignored second element"], 8, do_sort=True)   # first element only; do.sort no-op
    with pytest.raises(ValueError, match="not be longer than 31"):
        kh.seq_kmer_pos(ix, s, 32)
    with pytest.raises(ValueError, match="longer than k"):
        kh.seq_kmer_pos(ix, "ACGTACGT", 8)
    with pytest.raises(ValueError, match="single sequence"):
        kh.seq_kmer_pos(ix, [s, s], 8)
    r = kh.kmer_pos(ix, 8, canonical=True)
    assert r["kmer"] is None and r["pos"] is None and r["pair.pos"] is None and r["count"] is not None
    r = kh.kmer_pos(ix, 0, canonical=True)
    assert all(v is None for v in r.values())


def test_handles_coexist_and_free(kh, oracle):
    seqs = [random_dna(5000 + 100 * i, 40 + i, p_n=0.001) for i in range(5)]
    hs = [kh.make_kmer_hash(s, 9 + i) for i, s in enumerate(seqs)]
    for i in (3, 0, 4, 1, 2):
        o = oracle.build(seqs[i], 9 + i)
        assert np.array_equal(kh.kmer_pos(hs[i], 2, canonical=True)["pos"].ravel(), o.extract(2)["pos"])
    hs[2].free()
    hs[2].free()                         # finaliser tolerates a cleared pointer
    with pytest.raises(ValueError):
        kh.kmer_pos(hs[2], 2, canonical=True)


def test_pair_count_is_taken_on_demand(kh, oracle):
    """P (rows of pair.pos), the number of k-mers with pairs and the longest list are not part of the build: kmg_sizes with
    P = NULL never sweeps the index, the first request for P does (once), and the answer is the reference's."""
    import ctypes as C
    from kmer_hasher_b200 import _lib
    L = _lib.load()
    s = random_dna(60000, 77, p_n=0.002)
    s[1000:1400] = ord("T")                                # a k-mer with a few hundred positions
    for k, do_sort in ((12, False), (32, False), (21, True)):
        ix = kh.make_kmer_hash(s, k, do_sort=do_sort)
        o = oracle.build(s, k)
        kh.profile(enable=True, reset=True)
        assert ix.sizes_un == (o.U, o.N)
        got = kh.kmer_pos(ix, 2 | 8, canonical=True)       # neither asks for P
        assert "stats" not in kh.profile(reset=True)
        assert np.array_equal(got["count"], o.extract(8)["count"])
        assert ix.sizes == (o.U, o.N, o.P)                  # the sweep runs now ...
        assert kh.profile(reset=True)["stats"][1] == 1
        assert ix.sizes == (o.U, o.N, o.P)                  # ... and only once
        assert "stats" not in kh.profile(enable=False)
        kh.profile(reset=True)
        U, N, P = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _lib.check(L.kmg_sizes(ix._handle(), None, None, C.byref(P)))
        assert P.value == o.P
        assert np.array_equal(kh.kmer_pos(ix, 4, canonical=True)["pair.pos"].ravel(), o.extract(4)["pair_pos"])
        ix.free()


def test_pinned_and_device_buffers(kh, oracle):
    import torch
    s = random_dna(100000, 9, p_n=0.001)
    o = oracle.build(s, 21)
    # pinned host input + pinned outputs
    pin = kh.pinned_empty(len(s), np.uint8)
    pin[:] = s
    g = kh.make_kmer_hash(pin, 21, do_sort=True)
    out = {"pos": kh.pinned_empty((o.N, 2), np.int32), "count": kh.pinned_empty(o.U, np.int32)}
    e = kh.kmer_pos(g, 10, out=out)
    assert np.array_equal(e["pos"].ravel(), o.extract(2)["pos"])
    assert np.array_equal(e["count"], o.extract(8)["count"])
    # device-resident input and output
    d = torch.from_numpy(s).cuda()
    g2 = kh.make_kmer_hash(d, 21, do_sort=True)
    dout = torch.empty((o.N, 2), dtype=torch.int32, device="cuda")
    e2 = kh.kmer_pos(g2, 2, out={"pos": dout})
    torch.cuda.synchronize()
    assert np.array_equal(dout.cpu().numpy().ravel(), o.extract(2)["pos"])


@pytest.mark.parametrize("name,k", [("c2", 32), ("c3", 21), ("c5", 12)])
def test_scaled_configs(kh, oracle, name, k):
    """The BASELINE configurations at 1/20 scale, bit-exact against the oracle."""
    from kmer_hasher_b200 import synth
    s = {"c2": lambda: synth.config_c2(2_000_000), "c3": lambda: synth.config_c3(4_000_000, tail_k=21),
         "c5": lambda: synth.generate(1_000_000, 0xC5, tandem=0.01, tandem_unit_max=40, tandem_len_max=3000)}[name]()
    flags = 15 if name == "c5" else 1 | 2 | 8
    g, o = compare_index(kh, oracle, s, k, flags=flags)
    if name == "c3":
        q = synth.config_c4_query(s, 1_000_000)
        g32 = kh.make_kmer_hash(s, 32)
        o32 = oracle.build(s, 32)
        assert np.array_equal(kh.seq_kmer_pos(g32, q, 32, allow_k32=True).ravel(), o32.query(q, 32))


def test_full_size_properties(kh):
    """40 Mbp (config 2 size), k = 32: size-independent properties of the result."""
    from kmer_hasher_b200 import synth
    s = synth.config_c2(40_000_000)
    k = 32
    g = kh.make_kmer_hash(s, k)
    U, N, P = g.sizes
    assert N == len(s) - k + 1                       # no N in config 2
    e = kh.kmer_pos(g, 2 | 8)                        # the index's own (grouped) order: every property below holds in any order
    keys = kh.kmer_keys(g)
    assert len(np.unique(keys)) == U                 # distinct
    cnt = e["count"].astype(np.int64)
    assert cnt.sum() == N and cnt.min() >= 1
    assert P == int((cnt * (cnt - 1) // 2).sum())
    pos = e["pos"]
    assert np.array_equal(pos[:, 0], np.repeat(np.arange(1, U + 1, dtype=np.int32), cnt))
    p = pos[:, 1]
    assert np.array_equal(np.sort(p), np.arange(1, N + 1, dtype=np.int32))   # every window exactly once
    same = pos[1:, 0] == pos[:-1, 0]
    assert np.all(p[1:][same] > p[:-1][same])        # ascending inside each k-mer
    # every position's window re-encodes to its k-mer's key (checked on a sample)
    rng = np.random.default_rng(0)
    idx = rng.integers(0, N, 200000)
    code = ((s >> 1) & 3).astype(np.uint64)
    want = np.zeros(len(idx), np.uint64)
    starts = p[idx].astype(np.int64) - 1
    for j in range(k):
        want = (want << np.uint64(2)) | code[starts + j]
    assert np.array_equal(want, keys[pos[idx, 0] - 1])
    # self-probe of a 1 Mbp slice.  Repeats make the full result astronomically large (an R matrix could
    # not hold it), so count it, then stream and check two windows of rows through the chunk API.
    import ctypes as C
    from kmer_hasher_b200 import _lib
    L = _lib.load()
    q = np.ascontiguousarray(s[5_000_000:6_000_000])
    st, M = C.c_void_p(), C.c_uint64()
    _lib.check(L.kmg_query_begin(g._handle(), q.ctypes.data, len(q), k, C.byref(st), C.byref(M)))
    assert M.value >= len(q) - k + 1
    for first in (0, M.value // 2):
        n = min(2_000_000, M.value - first)
        rows = np.empty((n, 2), np.int32)
        _lib.check(L.kmg_query_emit_chunk(st, first, n, rows.ctypes.data))
        i0 = rows[:, 0].astype(np.int64) - k             # 0-based start in q
        j0 = rows[:, 1].astype(np.int64) - 1
        assert np.all(np.diff(rows[:, 0]) >= 0)
        same = rows[1:, 0] == rows[:-1, 0]
        assert np.all(rows[1:, 1][same] > rows[:-1, 1][same])
        sel = rng.integers(0, n, 100000)
        for j in (0, 7, 15, 31):
            assert np.array_equal(code[5_000_000 + i0[sel] + j], code[j0[sel] + j])
    L.kmg_query_free(st)


@pytest.mark.parametrize("world,k", [(1, 16), (2, 32), (4, 21), (8, 12)])
def test_sharded_engine_single_process(kh, oracle, world, k):
    """The CUDA side of the sharded build (kmg_shard_sample / kmg_shard_partition / kmg_build_records /
    kmg_query_records), with the ranks played one after another on one GPU: shard -> partition by
    splitters -> concatenate per owner in source order -> per-owner index == slices of the whole index."""
    import torch
    from kmer_hasher_b200 import dist as kdist, synth
    eng = kdist.CudaEngine(torch.device("cuda", 0))
    seq = synth.config_c3(600_000, tail_k=k)
    L = len(seq)
    per = (L + world - 1) // world
    if world > 1:
        seq[per - 2:per + 1] = np.frombuffer(b"nAC", np.uint8)      # breaker at the first cut
    whole = oracle.build(seq, k)
    want = whole.extract(2 | 8)
    shards, samples = [], []
    for r in range(world):
        s0, s1, g0, g1 = kdist.shard_bounds(L, world, r, k)
        t = eng.upload(seq[g0:g1])
        shards.append((t, g0, g1, s0, s1))
        samples.append(eng.sample(t, g0, g1, L, s0, s1, k, 512).cpu().numpy().view(np.uint64))
    spl = kdist.choose_splitters(np.concatenate(samples), world)
    per_owner_k = [[] for _ in range(world)]
    per_owner_p = [[] for _ in range(world)]
    for (t, g0, g1, s0, s1) in shards:
        keys, pos, counts = eng.partition(t, g0, g1, L, s0, s1, k, spl, world)
        assert sum(counts) <= s1 - s0
        off = 0
        for o, c in enumerate(counts):
            per_owner_k[o].append(keys[off:off + c])
            per_owner_p[o].append(pos[off:off + c])
            off += c
    got_keys, got_cnt, got_pos, offs = [], [], [], 0
    q = synth.config_c4_query(seq, 100_000)
    qo = whole.query(q, k)
    qkeys, qpos = oracle.windows(q, k)
    rows_all = []
    for o in range(world):
        kk = torch.cat(per_owner_k[o]); pp = torch.cat(per_owner_p[o])
        ix = eng.build_records(kk.clone(), pp.clone(), kk.numel(), k)
        U, N, _ = ix.sizes
        e = kh.kmer_pos(ix, 2 | 8, canonical=True)
        got_keys.append(kh.kmer_keys(ix, canonical=True)); got_cnt.append(e["count"])
        p = e["pos"].copy(); p[:, 0] += offs; got_pos.append(p)
        offs += U
        # routed probe: this owner's share of the query windows
        own = np.searchsorted(spl, qkeys, side="right") == o if world > 1 else np.ones(len(qkeys), bool)
        dk = torch.from_numpy(qkeys[own].view(np.int64).copy()).cuda()
        di = torch.from_numpy((qpos[own] + (k - 1)).astype(np.int32)).cuda()
        rows_all.append(eng.query_records(ix, dk, di, dk.numel()).cpu().numpy())
        ix.free()
    assert np.array_equal(np.concatenate(got_keys), want["keys"])
    assert np.array_equal(np.concatenate(got_cnt), want["count"])
    assert np.array_equal(np.concatenate(got_pos).ravel(), want["pos"])
    rows = np.concatenate(rows_all)
    rows = rows[np.argsort(rows[:, 0], kind="stable")]
    assert np.array_equal(rows.ravel(), qo)


@pytest.mark.parametrize("world,k,L,order", [(2, 32, 500_000, 1), (2, 32, 500_000, 0), (4, 27, 300_000, 0), (3, 21, 500_000, 0),
                                             (8, 12, 500_000, 1), (8, 32, 100, 1), (8, 32, 100, 0), (5, 7, 23, 0)])
def test_peer_scatter_single_process(kh, oracle, world, k, L, order):
    """The fused partition + exchange (kmg_shard_count / kmg_shard_scatter / kmg_build_received /
    kmg_query_received) with the ranks played one after another on one GPU: every "peer" array lives on
    this device, so the scatter's addressing (lower ranks first inside each owner's arrays), the
    device-side counts and the overflow flag are checked without NVLink."""
    import ctypes as C
    import torch
    from kmer_hasher_b200 import dist as kdist, synth, _lib
    L_ = _lib.load()
    dev = torch.device("cuda", 0)
    eng = kdist.CudaEngine(dev)
    seq = synth.config_c3(L, tail_k=k) if L > 1000 else synth.generate(L, 11, n_single=2)
    per = (L + world - 1) // world
    if L > 1000:
        seq[per - 2:per + 1] = np.frombuffer(b"nAC", np.uint8)      # breaker at the first cut
    whole = oracle.build(seq, k)
    want = whole.extract(2 | 8)
    # halo + splitter sample in one exchange (kmg_shard_pack -> "all-gather" -> kmg_shard_open_packed)
    ns = 512 if L > 1000 else 4
    owns = [eng.upload(seq[min(r * per, L):min((r + 1) * per, L)]) for r in range(world)]
    mixed = order == 0 and k >= 21                          # owners then hold ranges of the mixed key (grouped build: k >= 21)
    packs = [eng.shard_pack(o, k, ns, order) for o in owns]
    allpack = torch.cat(packs)
    handles, spls = [], []
    for r in range(world):
        h, spl_r = eng.shard_open_packed(owns[r], L, world, r, k, ns, allpack, order)
        handles.append(h); spls.append(spl_r.cpu().numpy().view(np.uint64))
    smp = np.concatenate([p.cpu().numpy()[48:].view(np.uint64) for p in packs])
    for r in range(world):
        assert np.all(np.diff(packs[r].cpu().numpy()[48:].view(np.uint64).astype(np.float64)) >= 0)
        assert np.array_equal(spls[r], kdist.choose_splitters(smp, world))
    spl = torch.from_numpy(spls[0].view(np.int64).copy()).to(dev)
    assert np.array_equal(kdist.device_splitters(torch.from_numpy(smp.view(np.int64).copy()).to(dev), world).cpu().numpy().view(np.uint64), spls[0])
    matrix = torch.cat([eng.shard_count(h, spl, world) for h in handles])
    m = matrix.cpu().numpy().reshape(world, world)
    assert m.sum() == whole.N
    cap = int(m.sum(axis=0).max()) + 7

    def scatter_all(cap, pos_add, hs, mat):
        keys = [torch.zeros(cap, dtype=torch.int64, device=dev) for _ in range(world)]
        pos = [torch.zeros(cap, dtype=torch.int32, device=dev) for _ in range(world)]
        infos = []
        for r in range(world):
            sl = kdist._Slot()
            sl.keys, sl.pos = keys[r].data_ptr(), pos[r].data_ptr()
            sl.peer_keys = (C.c_void_p * world)(*[t.data_ptr() for t in keys])
            sl.peer_pos = (C.c_void_p * world)(*[t.data_ptr() for t in pos])
            infos.append((sl, eng.shard_scatter(hs[r], spl, world, r, sl, cap, mat, pos_add)))
        return keys, pos, infos

    keys, pos, infos = scatter_all(cap, 0, handles, matrix)
    got_keys, got_cnt, got_pos, offs, owners = [], [], [], 0, []
    for o, (sl, info) in enumerate(infos):
        assert info.cpu().tolist() == [int(m[:, o].sum()), 0]
        ix = eng.build_received(sl, cap, info, k, order)
        U, N, _ = ix.sizes
        assert N == m[:, o].sum()
        assert (L_.kmg_index_order(ix._handle()) == 0) == mixed
        e = kh.kmer_pos(ix, 2 | 8)
        got_keys.append(kh.kmer_keys(ix)); got_cnt.append(e["count"])
        p = e["pos"].copy(); p[:, 0] += offs; got_pos.append(p)
        offs += U
        owners.append(ix)
    allk, allc, allp = np.concatenate(got_keys), np.concatenate(got_cnt), np.concatenate(got_pos)
    if mixed:                                               # owners hold ranges of the mix: order everything by key
        assert len(np.unique(allk)) == len(allk)
        order_k = np.argsort(allk, kind="stable")
        rank = np.empty(len(allk), np.int64); rank[order_k] = np.arange(len(allk))
        allk, allc = allk[order_k], allc[order_k]
        allp = kh._renumber(allp, rank)
    assert np.array_equal(allk, want["keys"])
    assert np.array_equal(allc, want["count"])
    assert np.array_equal(allp.ravel(), want["pos"])

    if L <= 1000:
        for ix in owners:
            ix.free()
        for h in handles:
            eng.shard_close(h)
        return
    # routed probe through the same scatter (coordinates = 1-based end of the query window)
    q = synth.config_c4_query(seq, 120_000)
    qo = whole.query(q, k)
    Lq = len(q)
    qh = []
    for r in range(world):
        s0, s1, g0, g1 = kdist.shard_bounds(Lq, world, r, k)
        qh.append(eng.shard_open(eng.upload(q[g0:g1]), g0, g1, Lq, s0, s1, k))
        if mixed:
            _lib.check(L_.kmg_shard_set_mixed(qh[-1], 1))
    qmat = torch.cat([eng.shard_count(h, spl, world) for h in qh])
    qcap = int(qmat.cpu().numpy().reshape(world, world).sum(axis=0).max()) + 1
    qkeep_k, qkeep_p, qinfos = scatter_all(qcap, k - 1, qh, qmat)      # keep the receive arrays alive
    rows = np.concatenate([eng.query_received(owners[o], sl, qcap, info, mixed).cpu().numpy() for o, (sl, info) in enumerate(qinfos)])
    rows = rows[np.lexsort((rows[:, 1], rows[:, 0]))]      # rows of one i may now come from... one owner still; j ascending
    assert np.array_equal(rows.ravel(), qo)

    # an exchange that is too small is reported, not silently truncated
    small = int(m.sum(axis=0).max()) - 1
    keep2_k, keep2_p, infos2 = scatter_all(small, 0, handles, matrix)
    big = int(np.argmax(m.sum(axis=0)))
    assert infos2[big][1].cpu().tolist()[1] == 1
    with pytest.raises(_lib.KmgError):
        eng.build_received(infos2[big][0], small, infos2[big][1], k, order)
    for ix in owners:
        ix.free()
    for h in handles + qh:
        eng.shard_close(h)


@pytest.mark.parametrize("ka,kb", [(16, 16), (32, 32), (12, 12), (9, 7)])
def test_kmer_pairs(kh, oracle, ka, kb):
    """kmer.pairs (kmg_join_*) against the restated kmer_pair_pos: shared k-mers' position pairs, a outer / b inner."""
    import oracle as oracle_mod
    from kmer_hasher_b200 import synth
    a_seq = synth.config_c3(300_000, tail_k=ka)
    b_seq = synth.config_c4_query(a_seq, 150_000)
    ia, ib = kh.make_kmer_hash(a_seq, ka, do_sort=True), kh.make_kmer_hash(b_seq, kb)   # rows follow a's k-mer order
    oa, ob = oracle.build(a_seq, ka), oracle.build(b_seq, kb)
    want = oracle_mod.pairs_join(oa.extract(2 | 8), ob.extract(2 | 8))
    got = kh.kmer_pairs(ia, ib)
    assert got.shape == (len(want) // 2, 2)
    assert np.array_equal(got.ravel(), want)
    ig = kh.make_kmer_hash(a_seq, ka)                     # grouped order of a: the same rows, k-mers in another order
    gg = kh.kmer_pairs(ig, ib)
    assert np.array_equal(gg[np.lexsort((gg[:, 1], gg[:, 0]))], got[np.lexsort((got[:, 1], got[:, 0]))])
    ig.free()
    # chunked emission == whole, pinned output, and an index joined with itself = sum of squares of the counts
    import ctypes as C
    from kmer_hasher_b200 import _lib
    L = _lib.load()
    st, M = C.c_void_p(), C.c_uint64()
    _lib.check(L.kmg_join_begin(ia._handle(), ib._handle(), C.byref(st), C.byref(M)))
    assert M.value == len(got)
    if M.value > 10:
        cut = M.value // 3
        p1, p2 = np.empty((cut, 2), np.int32), kh.pinned_empty((M.value - cut, 2), np.int32)
        _lib.check(L.kmg_join_emit_chunk(st, 0, cut, p1.ctypes.data))
        _lib.check(L.kmg_join_emit_chunk(st, cut, M.value - cut, p2.ctypes.data))
        assert np.array_equal(np.concatenate([p1, np.asarray(p2)]), got)
    L.kmg_join_free(st)
    if ka == kb:
        self_rows = kh.kmer_pairs(ia, ia)
        cnt = kh.kmer_pos(ia, 8, canonical=True)["count"].astype(np.int64)
        assert len(self_rows) == int((cnt * cnt).sum())
    empty = kh.make_kmer_hash("A" * 40, 16)
    other = kh.make_kmer_hash("C" * 40, 16)
    assert kh.kmer_pairs(empty, other).shape == (0, 2)
    for ix in (ia, ib, empty, other):
        ix.free()


def _revcomp_host(seq: np.ndarray) -> np.ndarray:
    """reverseComplement as Biostrings defines it for DNA: IUPAC complement, case kept, anything else unchanged."""
    tab = np.arange(256, dtype=np.uint8)
    for a, b in ("AT", "CG", "RY", "KM", "BV", "DH"):
        for x, y in ((a, b), (b, a)):
            tab[ord(x)] = ord(y)
            tab[ord(x.lower())] = ord(y.lower())
    return tab[seq[::-1]].copy()


@pytest.mark.parametrize("k", [4, 16, 31])
def test_reverse_complement_probe(kh, oracle, k):
    """seq.kmer.pos on the reverse strand with the reverse complement taken on the device == the reference's
    probe of the host-made reverse complement (what test.R:43-52 does with Biostrings)."""
    from kmer_hasher_b200 import synth
    seq = synth.config_c3(200_000, tail_k=k)
    q = synth.config_c4_query(seq, 60_000)
    q[100:110] = np.frombuffer(b"RYKMBVDHSW", np.uint8)          # IUPAC codes
    q[200:204] = np.frombuffer(b"rykm", np.uint8)
    q[300] = ord("-")
    rc = _revcomp_host(q)
    index_of_rc = kh.make_kmer_hash(_revcomp_host(seq), k)        # so that the reverse strand of q has hits
    o = oracle.build(_revcomp_host(seq), k)
    want = o.query(rc, k)
    got = kh.seq_kmer_pos(index_of_rc, q, k, reverse_complement=True)
    assert len(want) > 0 and np.array_equal(got.ravel(), want)
    assert np.array_equal(kh.seq_kmer_pos(index_of_rc, rc, k).ravel(), want)
    index_of_rc.free()


@pytest.mark.parametrize("rank,shape,rb", [(0, 0, 8), (3, 0, 8), (0, 0, 9), (3, 0, 9), (3, 1, 8), (3, 1, 9), (0, 1, 9),
                                           (3, 2, 8), (0, 2, 8), (3, 0, 10), (0, 0, 10)])
def test_every_sort_pass_variant_is_exact(kh, oracle, rank, shape, rb):
    """The pass variants selectable with kmg_tune -- rank by one atomic per record or by bitmap match, the tile
    shapes, digits of 8, 9 and 10 bits -- all give the reference's index; the one-atomic variant is only the default
    where the lane-order self-test passes, which is asserted here for this GPU."""
    import ctypes as C
    from kmer_hasher_b200 import synth, _lib
    L = _lib.load()
    f = C.c_uint32(1)
    _lib.check(L.kmg_selftest_lane_order(C.byref(f)))
    assert f.value == 0
    seq = synth.config_c2(400_000)
    seq[1000:3000] = ord("A")                               # a homopolymer: every lane of a warp on one bin
    seq[5000:9000] = np.frombuffer(b"CA" * 2000, np.uint8)
    seq[20000:20040] = ord("N")
    try:
        _lib.check(L.kmg_tune(b"sort_cfg", rank))
        _lib.check(L.kmg_tune(b"sort_shape", shape))
        _lib.check(L.kmg_tune(b"hash_rb", rb))
        for k in (32, 21, 9):
            ix = kh.make_kmer_hash(seq, k)
            got = kh.kmer_pos(ix, 2 | 8, canonical=True)
            want = oracle.build(seq, k).extract(2 | 8)
            assert np.array_equal(kh.kmer_keys(ix, canonical=True), want["keys"])
            assert np.array_equal(got["count"], want["count"])
            assert np.array_equal(got["pos"].ravel(), want["pos"])
            ix.free()
        assert L.kmg_tune_get(b"unstable_rebuilds", 0) == 0
    finally:
        _lib.check(L.kmg_tune(b"sort_cfg", -1))
        _lib.check(L.kmg_tune(b"sort_shape", -1))
        _lib.check(L.kmg_tune(b"hash_rb", 0))


def test_unstable_sort_is_caught_and_rebuilt(kh, oracle):
    """The always-on guard (rle_kernel checks that positions ascend inside every k-mer): a sort pass that is unstable on
    purpose (rank variant 4 visits a thread's records in reverse) is noticed, the device is switched to the bitmap
    variant for good and the index is rebuilt, so the caller still gets the reference's result."""
    import ctypes as C
    from kmer_hasher_b200 import synth, _lib
    L = _lib.load()
    seq = synth.config_c2(300_000)
    seq[1000:3000] = ord("A")
    seq[5000:9000] = np.frombuffer(b"CA" * 2000, np.uint8)
    before = L.kmg_tune_get(b"unstable_rebuilds", 0)
    try:
        for k, order in ((32, False), (12, True)):          # grouped and sorted builds
            _lib.check(L.kmg_tune(b"reset_rank", 0))
            _lib.check(L.kmg_tune(b"sort_cfg", 4))
            ix = kh.make_kmer_hash(seq, k, do_sort=order)
            assert L.kmg_tune_get(b"rank_variant", 0) == 0     # demoted to the bitmap variant, override dropped
            got = kh.kmer_pos(ix, 2 | 8, canonical=True)
            want = oracle.build(seq, k).extract(2 | 8)
            assert np.array_equal(got["count"], want["count"]) and np.array_equal(got["pos"].ravel(), want["pos"])
            ix.free()
        assert L.kmg_tune_get(b"unstable_rebuilds", 0) == before + 2
        # builds from records report it instead (the caller's arrays are consumed): KMG_ERR_UNSTABLE
        import torch
        okeys, opos = oracle.windows(seq, 32)
        keys = torch.from_numpy(okeys.view(np.int64)).cuda()
        pos = torch.from_numpy(opos).cuda()
        _lib.check(L.kmg_tune(b"reset_rank", 0))
        _lib.check(L.kmg_tune(b"sort_cfg", 4))
        h = C.c_void_p()
        rc = L.kmg_build_records(keys.data_ptr(), pos.data_ptr(), len(okeys), 32, C.byref(h))
        assert rc == -7 and L.kmg_tune_get(b"rank_variant", 0) == 0
    finally:
        _lib.check(L.kmg_tune(b"sort_cfg", -1))
        _lib.check(L.kmg_tune(b"reset_rank", 0))              # the self-test decides again
    assert L.kmg_tune_get(b"rank_variant", 0) == 3


@pytest.mark.parametrize("bits", [16, 24, 27, 40])
def test_grouped_build_with_forced_collisions(kh, oracle, bits):
    """KMG_ORDER_GROUPED sorts on `bits` bits of a mix of the key and partitions the groups in which several
    k-mers share those bits.  Few bits force such groups everywhere: 8 bits -> every group is long (the
    block-level partition), 16 bits -> every group is short (the in-place insertion sort), 24 -> a mixture."""
    import ctypes as C
    from kmer_hasher_b200 import synth, _lib
    L = _lib.load()
    seq = synth.generate(200_000, 0xC2, repeat=0.3, lower=0.2)
    seq[1000:1900] = ord("A")                              # a k-mer with ~870 positions: its group is a long one
    seq[9000:10200] = np.frombuffer(b"CA" * 600, np.uint8)
    seq[50000:50003] = ord("n")
    try:
        _lib.check(L.kmg_tune(b"hash_bits", bits))
        for k in (32, 27):
            if bits // 8 + 1 >= (2 * k + 7) // 8:          # not worth grouping: the build is sorted anyway
                continue
            ix = kh.make_kmer_hash(seq, k)
            assert L.kmg_index_order(ix._handle()) == 0
            o = oracle.build(seq, k)
            want = o.extract(15)
            assert ix.sizes == (o.U, o.N, o.P)
            got = kh.kmer_pos(ix, 15, canonical=True)
            assert np.array_equal(kh.kmer_keys(ix, canonical=True), want["keys"])
            assert np.array_equal(got["count"], want["count"])
            assert np.array_equal(got["pos"].ravel(), want["pos"])
            assert np.array_equal(got["pair.pos"].ravel(), want["pair_pos"])
            raw = kh.kmer_pos(ix, 2 | 8)                   # in the index's own order: rows grouped by i, lists ascending
            assert np.array_equal(raw["pos"][:, 0], np.repeat(np.arange(1, o.U + 1, dtype=np.int32), raw["count"]))
            q = synth.config_c4_query(seq, 80_000)
            assert np.array_equal(kh.seq_kmer_pos(ix, q, k, allow_k32=True).ravel(), o.query(q, k))
            srt = kh.make_kmer_hash(seq, k, do_sort=True)  # the sorted build gives ascending keys directly
            assert L.kmg_index_order(srt._handle()) == 1
            assert np.array_equal(kh.kmer_keys(srt), want["keys"])
            ix.free(); srt.free()
        # exact copies of one block: every k-mer has three positions, so k-mers that share the sorted bits always
        # interleave (A B A B A B) -- the groups that do get filed and fixed; above, most colliding k-mers have one
        # position each and are left as they lie
        if bits <= 27:
            seq3 = np.tile(synth.generate(70_000, 9, lower=0.1), 3)
            ix = kh.make_kmer_hash(seq3, 32)
            o = oracle.build(seq3, 32)
            want = o.extract(2 | 8)
            assert ix.sizes_un == (o.U, o.N)
            got = kh.kmer_pos(ix, 2 | 8, canonical=True)
            assert np.array_equal(kh.kmer_keys(ix, canonical=True), want["keys"])
            assert np.array_equal(got["count"], want["count"]) and np.array_equal(got["pos"].ravel(), want["pos"])
            q = synth.config_c4_query(seq3, 50_000)
            assert np.array_equal(kh.seq_kmer_pos(ix, q, 32, allow_k32=True).ravel(), o.query(q, 32))
            ix.free()
    finally:
        _lib.check(L.kmg_tune(b"hash_bits", 0))


def test_grouped_build_falls_back_when_the_task_list_overflows(kh, oracle):
    """More colliding groups than the fix-up's task list holds (forced: 16 hash bits, a 16-entry list): the
    build notices and redoes itself sorted by key; the sharded owner build reports the condition instead."""
    import ctypes as C
    import torch
    from kmer_hasher_b200 import synth, _lib, dist as kdist
    L = _lib.load()
    # three exact copies of one block: every k-mer has three positions, so wherever two k-mers share the sorted bits their
    # records interleave (only such groups are filed: k-mers that collide but already sit one after the other are left alone)
    seq = np.tile(synth.generate(33_000, 5), 3)
    k = 32
    try:
        _lib.check(L.kmg_tune(b"hash_bits", 16))
        _lib.check(L.kmg_tune(b"fix_cap", 16))
        ix = kh.make_kmer_hash(seq, k)
        assert L.kmg_index_order(ix._handle()) == 1          # fell back to the sorted build
        want = oracle.build(seq, k).extract(2 | 8)
        got = kh.kmer_pos(ix, 2 | 8)
        assert np.array_equal(kh.kmer_keys(ix), want["keys"])
        assert np.array_equal(got["count"], want["count"]) and np.array_equal(got["pos"].ravel(), want["pos"])
        ix.free()
        # owner-side build of received records: refused with KMG_ERR_RANGE (dist.py then takes the general path)
        eng = kdist.CudaEngine(torch.device("cuda", 0))
        own = eng.upload(seq)
        pack = eng.shard_pack(own, k, 64, 0)
        h, spl = eng.shard_open_packed(own, len(seq), 1, 0, k, 64, pack, 0)
        counts = eng.shard_count(h, spl, 1)
        cap = len(seq)
        keys = torch.zeros(cap, dtype=torch.int64, device="cuda"); pos = torch.zeros(cap, dtype=torch.int32, device="cuda")
        sl = kdist._Slot()
        sl.keys, sl.pos = keys.data_ptr(), pos.data_ptr()
        sl.peer_keys, sl.peer_pos = (C.c_void_p * 1)(keys.data_ptr()), (C.c_void_p * 1)(pos.data_ptr())
        info = eng.shard_scatter(h, spl, 1, 0, sl, cap, counts, 0)
        with pytest.raises(_lib.KmgError) as e:
            eng.build_received(sl, cap, info, k, 0)
        assert e.value.code == -3
        eng.shard_close(h)
    finally:
        _lib.check(L.kmg_tune(b"hash_bits", 0))
        _lib.check(L.kmg_tune(b"fix_cap", 0))


@pytest.mark.parametrize("world,k,L", [(2, 32, 500_000), (4, 27, 300_000), (3, 21, 500_000), (8, 32, 2_000_000), (8, 32, 100), (5, 21, 30)])
def test_region_exchange_single_process(kh, oracle, world, k, L):
    """The region exchange of the grouped sharded build (kmg_shard_scatter_ranges / kmg_build_regions /
    kmg_query_regions) with the ranks played one after another on one GPU: owners are equal ranges of the mixed key,
    every source writes into its own region of every owner's arrays, the scatter's last tile leaves the counts with
    the owners.  The owners' indexes together must be the reference's index; sharded extraction numbers k-mers globally."""
    import ctypes as C
    import torch
    from kmer_hasher_b200 import dist as kdist, synth, _lib
    L_ = _lib.load()
    dev = torch.device("cuda", 0)
    eng = kdist.CudaEngine(dev)
    seq = synth.config_c3(L, tail_k=k) if L > 1000 else synth.generate(L, 11, n_single=2)
    per = (L + world - 1) // world
    if L > 1000:
        seq[per - 2:per + 1] = np.frombuffer(b"nAC", np.uint8)      # breaker at the first cut
        seq[3000:3000 + 40_000 // world] = ord("A")                  # a homopolymer: one k-mer, one owner, many copies
    whole = oracle.build(seq, k)
    want = whole.extract(2 | 8)
    owns = [eng.upload(seq[min(r * per, L):min((r + 1) * per, L)]) for r in range(world)]
    packs = [eng.shard_pack(o, k, 2, 0) for o in owns]
    allpack = torch.cat(packs)
    handles = [eng.shard_open_packed(owns[r], L, world, r, k, 2, allpack, 0, splitters=False)[0] for r in range(world)]

    def exchange(hs, region_cap, pos_add):
        cap = region_cap * world
        keys = [torch.zeros(cap, dtype=torch.int64, device=dev) for _ in range(world)]
        pos = [torch.zeros(cap, dtype=torch.int32, device=dev) for _ in range(world)]
        cnts = [torch.full((16,), -1, dtype=torch.int64, device=dev) for _ in range(world)]
        slots = []
        for r in range(world):
            sl = kdist._Slot()
            sl.keys, sl.pos, sl.counts = keys[r].data_ptr(), pos[r].data_ptr(), cnts[r].data_ptr()
            sl.peer_keys = (C.c_void_p * world)(*[t.data_ptr() for t in keys])
            sl.peer_pos = (C.c_void_p * world)(*[t.data_ptr() for t in pos])
            sl.peer_counts = (C.c_void_p * world)(*[t.data_ptr() for t in cnts])
            slots.append(sl)
            eng.shard_scatter_ranges(hs[r], world, r, sl, region_cap, pos_add)
        torch.cuda.synchronize()
        return slots, (keys, pos, cnts)

    region_cap = int(per / world * 1.5) + 50_000
    slots, keep = exchange(handles, region_cap, 0)
    sent = torch.stack(keep[2])[:, :world].cpu().numpy()            # [owner][source]
    assert sent.min() >= 0 and sent.sum() == whole.N
    owners, U_all, N_all = [], [], []
    for o in range(world):
        ix = eng.build_regions(slots[o], region_cap, world, k)
        assert L_.kmg_index_order(ix._handle()) == 0 and ix.sizes[1] == sent[o].sum()
        owners.append(ix); U_all.append(ix.sizes[0]); N_all.append(ix.sizes[1])
    assert sum(U_all) == whole.U and sum(N_all) == whole.N
    # sharded extraction: every owner writes its slice of ONE matrix with global k-mer numbers
    pos_all = torch.zeros((whole.N, 2), dtype=torch.int32, device=dev)
    cnt_all = torch.zeros(whole.U, dtype=torch.int32, device=dev)
    keys_all = np.empty(whole.U, np.uint64)
    for o, ix in enumerate(owners):
        sh = kdist.ShardedIndex(ix, k, o, world, U_all, N_all, None, eng)
        r0, i0 = sh.row_offset, sh.i_offset
        got = sh.kmer_pos(2 | 8, out={"pos": pos_all[r0:r0 + N_all[o]], "count": cnt_all[i0:i0 + U_all[o]]})
        assert got["i_offset"] == i0 and got["row_offset"] == r0
        keys_all[i0:i0 + U_all[o]] = kh.kmer_keys(ix)
    pos_all, cnt_all = pos_all.cpu().numpy(), cnt_all.cpu().numpy()
    assert np.array_equal(pos_all[:, 0], np.repeat(np.arange(1, whole.U + 1, dtype=np.int32), cnt_all))   # global i, grouped
    order = np.argsort(keys_all, kind="stable")
    rank = np.empty(whole.U, np.int64); rank[order] = np.arange(whole.U)
    assert np.array_equal(keys_all[order], want["keys"]) and np.array_equal(cnt_all[order], want["count"])
    assert np.array_equal(kh._renumber(pos_all, rank).ravel(), want["pos"])
    # routed probe: query records go to the key's owner through the same scatter (pos_add = k - 1: the END coordinate)
    if L > 1000:
        q = synth.config_c4_query(seq, L // 3)
        qper = (len(q) + world - 1) // world
        qowns = [eng.upload(q[min(r * qper, len(q)):min((r + 1) * qper, len(q))]) for r in range(world)]
        qpacks = torch.cat([eng.shard_pack(o, k, 2, 0) for o in qowns])
        qh = [eng.shard_open_packed(qowns[r], len(q), world, r, k, 2, qpacks, 0, splitters=False)[0] for r in range(world)]
        qslots, qkeep = exchange(qh, int(qper / world * 1.5) + 50_000, k - 1)
        rows = [eng.query_regions(owners[o], qslots[o], int(qper / world * 1.5) + 50_000, world).cpu().numpy() for o in range(world)]
        rows = np.concatenate(rows)
        rows = rows[np.argsort(rows[:, 0], kind="stable")]
        assert len(rows) > 0 and np.array_equal(rows.ravel(), whole.query(q, k))
        for h in qh:
            eng.shard_close(h)
        # a region too small for what a source sends is reported, not overrun
        tiny = max(per // world // 4, 1)
        tslots, tkeep = exchange(handles, tiny, 0)
        with pytest.raises(_lib.KmgError) as e:
            eng.build_regions(tslots[0], tiny, world, k)
        assert e.value.code == -3
    for ix in owners:
        ix.free()
    for h in handles:
        eng.shard_close(h)


def test_arena_cache_can_be_trimmed(kh):
    """an R session must be able to give the cached device memory of freed indexes back (kmg_trim)."""
    from kmer_hasher_b200 import synth, _lib
    L_ = _lib.load()
    seq = synth.config_c2(2_000_000)
    kh.make_kmer_hash(seq, 32).free()
    assert L_.kmg_cached_bytes() > 0                       # freed blocks are kept for the next build ...
    _lib.check(L_.kmg_trim())
    assert L_.kmg_cached_bytes() == 0                      # ... until asked to return them
    ix = kh.make_kmer_hash(seq, 32)                        # and the library keeps working
    assert ix.sizes[1] == len(seq) - 31
    ix.free()


def test_pageable_buffers_take_the_staged_path(kh, oracle):
    """Large pageable inputs/outputs (what R hands the glue) go through pinned staging slots and host copy threads; the
    result must be the same bytes as with device or pinned buffers (chunk boundaries, odd sizes)."""
    import torch
    from kmer_hasher_b200 import synth
    seq = synth.config_c3(12_000_003)                      # > 8 MB: staged upload; pos matrix 96 MB: two staged chunks
    ix_pageable = kh.make_kmer_hash(seq, 21)               # numpy array = pageable
    ix_device = kh.make_kmer_hash(torch.from_numpy(seq).cuda(), 21)
    assert ix_pageable.sizes == ix_device.sizes
    U, N, _ = ix_device.sizes
    got = kh.kmer_pos(ix_pageable, 2 | 8)                  # pageable numpy outputs
    dpos = torch.empty((N, 2), dtype=torch.int32, device="cuda")
    dcnt = torch.empty(U, dtype=torch.int32, device="cuda")
    kh.kmer_pos(ix_device, 2 | 8, out={"pos": dpos, "count": dcnt})
    assert np.array_equal(got["pos"], dpos.cpu().numpy()) and np.array_equal(got["count"], dcnt.cpu().numpy())
    pin = kh.pinned_empty((N, 2), np.int32)
    kh.kmer_pos(ix_device, 2, out={"pos": pin})
    assert np.array_equal(pin, got["pos"])
    ix_pageable.free(); ix_device.free()


def test_first_pass_regions_and_their_overflow_fallback(kh, oracle):
    """The grouped build's first pass can write every bin of the mixed digit into its own region of 1.25x the mean bin
    size (no histogram sweep over the sequence).  A 1-in-251 sample decides whether the heaviest bin stays clear of a
    region; if a bin outgrows its region all the same (forced here), the build notices -- device flag, the reading pass
    is handed an empty source -- and redoes itself with the histogram.  Same index every way."""
    from kmer_hasher_b200 import synth, _lib
    L_ = _lib.load()
    k = 32
    seq = synth.config_c3(3_000_000)
    want = oracle.build(seq, k).extract(2 | 8)

    def check(sq, wanted):
        ix = kh.make_kmer_hash(sq, k)
        got = kh.kmer_pos(ix, 2 | 8, canonical=True)
        assert np.array_equal(kh.kmer_keys(ix, canonical=True), wanted["keys"])
        assert np.array_equal(got["count"], wanted["count"]) and np.array_equal(got["pos"].ravel(), wanted["pos"])
        ix.free()
    before = L_.kmg_tune_get(b"region_rebuilds", 0)
    try:
        for mode in (0, 1, 2):                               # sampled choice, histogram always, regions always
            _lib.check(L_.kmg_tune(b"no_regions", mode))
            check(seq, want)
        assert L_.kmg_tune_get(b"region_rebuilds", 0) == before
        heavy = seq.copy()
        heavy[100_000:160_000] = ord("A")                    # one k-mer with 59,969 copies: mean bin 11.7 k, region 18.7 k
        heavy[300_000:340_000] = np.frombuffer(b"CA" * 20_000, np.uint8)
        want_heavy = oracle.build(heavy, k).extract(2 | 8)
        _lib.check(L_.kmg_tune(b"no_regions", 0))
        check(heavy, want_heavy)                             # the sample sees the heavy bin: histogram path, nothing redone
        assert L_.kmg_tune_get(b"region_rebuilds", 0) == before
        _lib.check(L_.kmg_tune(b"no_regions", 2))
        check(heavy, want_heavy)                             # forced: the bin overflows, the build is redone
        assert L_.kmg_tune_get(b"region_rebuilds", 0) == before + 1
    finally:
        _lib.check(L_.kmg_tune(b"no_regions", 0))
