"""Full-size runs of the BASELINE configurations C3/C4/C5 (not part of pytest: minutes of host work).
usage: python tools/fullscale.py [c3] [c4] [c5] [ref]   ('ref' adds the 250 Mbp parity check against the
reference engine: ~3 min and ~35 GB of host memory)"""
import ctypes as C
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import kmer_hasher_b200 as kh
from kmer_hasher_b200 import _lib, synth

L = _lib.load()
what = set(sys.argv[1:]) or {"c3", "c4", "c5"}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) * 1e3


def show_profile(tag):
    for n, v in sorted(kh.profile(reset=True).items()):
        if v[0] > 0:
            print(f"    [{tag}] {n:18s} {v[0]:8.3f} ms {v[1]:3d} launches {v[2] / v[0] / 1e6 if v[2] else 0:7.0f} GB/s")


kh.profile(enable=True, reset=True)
if what & {"c3", "c4", "ref", "ref32"}:
    t0 = time.time()
    seq = synth.config_c3(250_000_000)
    print(f"C3 sequence generated in {time.time() - t0:.1f}s; N bytes {(seq | 0x20 == ord('n')).sum()}")
    dseq = torch.from_numpy(seq).cuda()

if "c3" in what:
    kh.make_kmer_hash(dseq, 21).free()
    kh.profile(reset=True)
    ix, ms = timed(lambda: kh.make_kmer_hash(dseq, 21))
    U, N, P = ix.sizes
    print(f"C3 build k=21: {ms:.2f} ms, N={N} U={U} P={P} -> {N / ms / 1e6:.2f} G k-mers/s")
    show_profile("c3 build")
    pos = torch.empty((N, 2), dtype=torch.int32, device="cuda")
    cnt = torch.empty(U, dtype=torch.int32, device="cuda")
    _, ms = timed(lambda: kh.kmer_pos(ix, 10, out={"pos": pos, "count": cnt}))
    print(f"C3 kmer.pos(2|8) device-resident: {ms:.2f} ms")
    show_profile("c3 extract")
    c = cnt.long()
    assert int(c.sum()) == N and int((c * (c - 1) // 2).sum()) == P
    p = pos[:, 1]
    srt = torch.sort(p).values
    assert bool((srt[1:] > srt[:-1]).all())                      # every window at most once
    same = pos[1:, 0] == pos[:-1, 0]
    assert bool((p[1:][same] > p[:-1][same]).all())              # ascending inside each k-mer
    keys = torch.from_numpy(kh.kmer_keys(ix).view(np.int64)).cuda()
    # every listed position is an N-free window (checked on a sample) and re-encodes to its key
    idx = torch.randint(0, N, (2_000_000,), device="cuda")
    st = p[idx].long() - 1
    code = ((dseq >> 1) & 3).long()
    w = torch.zeros_like(st)
    for j in range(21):
        w = (w << 2) | code[st + j]
        assert not bool(((dseq[st + j] | 0x20) == ord("n")).any())
    assert bool((w == keys[pos[idx, 0].long() - 1]).all())
    print("C3 properties ok; sha(count)=%s sha(pos)=%s" % (sha(cnt.cpu().numpy()), sha(pos.cpu().numpy())))
    if "ref" in what:
        from oracle import Reference
        t0 = time.time()
        r = Reference().build(seq, 21)
        print(f"reference seq_to_hash 250 Mbp k=21: {r.build_seconds:.1f}s -> {r.N / r.build_seconds / 1e6:.2f} M k-mers/s")
        e = r.extract(2 | 8)
        print(f"reference canonical extraction took {time.time() - t0 - r.build_seconds:.1f}s")
        assert (r.U, r.N, r.P) == (U, N, P)
        assert np.array_equal(e["count"], cnt.cpu().numpy())
        assert np.array_equal(e["pos"], pos.cpu().numpy().ravel())
        assert np.array_equal(e["keys"], kh.kmer_keys(ix))
        print("C3 FULL-SIZE PARITY with the reference engine: keys, counts, (i,pos) identical")
        r.close()
    del pos, cnt
    ix.free()

if "ref32" in what:
    # grouped build at scale against the reference engine: 100 Mbp, k=32 (thousands of colliding 40-bit groups)
    from oracle import Reference
    Lr = 100_000_000
    sub = np.ascontiguousarray(seq[:Lr])
    ix, ms = timed(lambda: kh.make_kmer_hash(torch.from_numpy(sub).cuda(), 32))
    U, N, P = ix.sizes
    assert L.kmg_index_order(ix._handle()) == 0
    t0 = time.time()
    r = Reference().build(sub, 32)
    print(f"ref32: reference seq_to_hash 100 Mbp k=32: {r.build_seconds:.1f}s -> {r.N / r.build_seconds / 1e6:.2f} M k-mers/s; ours (grouped) {ms:.2f} ms")
    e = r.extract(2 | 8)
    assert (r.U, r.N, r.P) == (U, N, P)
    got = kh.kmer_pos(ix, 2 | 8, canonical=True)
    assert np.array_equal(kh.kmer_keys(ix, canonical=True), e["keys"])
    assert np.array_equal(got["count"], e["count"])
    assert np.array_equal(got["pos"].ravel(), e["pos"])
    print(f"ref32 FULL-SIZE PARITY of the grouped build with the reference engine (after ordering by k-mer): keys, counts, (i,pos) identical; took {time.time() - t0:.0f}s")
    r.close()
    ix.free()

if "c4" in what:
    Lq = 100_000_000
    q = synth.config_c4_query(seq, Lq)
    dq = torch.from_numpy(q).cuda()
    ix, ms = timed(lambda: kh.make_kmer_hash(dseq, 32))
    U, N, P = ix.sizes
    print(f"C4 index k=32: {ms:.2f} ms, N={N} U={U}")
    kh.profile(reset=True)
    st, M = C.c_void_p(), C.c_uint64()
    for rep in range(2):
        _, ms = timed(lambda: _lib.check(L.kmg_query_begin(ix._handle(), dq.data_ptr(), Lq, 32, C.byref(st), C.byref(M))))
        if rep == 0:
            L.kmg_query_free(st)
    print(f"C4 probe (encode+match+scan) {ms:.2f} ms -> {(Lq - 31) / ms / 1e6:.2f} G k-mers/s queried; M={M.value} rows")
    show_profile("c4 match")
    if M.value < 2**31:
        rows = torch.empty((M.value, 2), dtype=torch.int32, device="cuda")
        _, ms = timed(lambda: _lib.check(L.kmg_query_emit(st, rows.data_ptr())))
        print(f"C4 emit {ms:.2f} ms -> {M.value / ms / 1e6:.2f} G rows/s ({8 * M.value / ms / 1e6:.0f} GB/s written)")
        show_profile("c4 emit")
        i0 = rows[:, 0].long() - 32
        j0 = rows[:, 1].long() - 1
        assert bool((rows[1:, 0] >= rows[:-1, 0]).all())
        same = rows[1:, 0] == rows[:-1, 0]
        assert bool((rows[1:, 1][same] > rows[:-1, 1][same]).all())
        code_i, code_q = ((dseq >> 1) & 3), ((dq >> 1) & 3)
        sel = torch.randint(0, M.value, (1_000_000,), device="cuda")
        for j in range(32):
            assert bool((code_q[i0[sel] + j] == code_i[j0[sel] + j]).all())
        print("C4 properties ok; sha(rows)=%s" % sha(rows.cpu().numpy()))
    else:
        print("C4: M exceeds an R matrix; streamed chunk check only")
    L.kmg_query_free(st)
    ix.free()

if "c5" in what:
    s5 = synth.config_c5()
    d5 = torch.from_numpy(s5).cuda()
    ix, ms = timed(lambda: kh.make_kmer_hash(d5, 12))
    U, N, P = ix.sizes
    print(f"C5 build k=12: {ms:.2f} ms, N={N} U={U} P={P} (target 1e9 < P < 2^31-1: {1e9 < P < 2**31 - 1})")
    if P < 2**31:
        kh.profile(reset=True)
        out = kh.pinned_empty((P, 3), np.int32)
        _, ms = timed(lambda: _lib.check(L.kmg_pairs(ix._handle(), out.ctypes.data)))
        _, ms = timed(lambda: _lib.check(L.kmg_pairs(ix._handle(), out.ctypes.data)))
        print(f"C5 pair.pos streamed to pinned host: {ms:.1f} ms -> {P / ms / 1e6:.2f} G pairs/s, {12 * P / ms / 1e6:.1f} GB/s over PCIe")
        show_profile("c5 pairs")
        dev = torch.empty((P, 3), dtype=torch.int32, device="cuda")
        _, ms = timed(lambda: _lib.check(L.kmg_pairs(ix._handle(), dev.data_ptr())))
        print(f"C5 pair.pos device-resident: {ms:.1f} ms -> {12 * P / ms / 1e6:.0f} GB/s written")
        show_profile("c5 pairs dev")
        assert np.array_equal(out[:5_000_000], dev[:5_000_000].cpu().numpy()) and np.array_equal(out[-5_000_000:], dev[-5_000_000:].cpu().numpy())
        assert bool((dev[:, 1] < dev[:, 2]).all()) and bool((dev[1:, 0] >= dev[:-1, 0]).all())
        cnt = torch.from_numpy(kh.kmer_pos(ix, 8)["count"]).cuda().long()
        per = torch.bincount(dev[:, 0].long() - 1, minlength=U)
        assert bool((per == cnt * (cnt - 1) // 2).all())
        code = ((d5 >> 1) & 3)
        sel = torch.randint(0, P, (1_000_000,), device="cuda")
        for j in range(12):
            assert bool((code[dev[sel, 1].long() - 1 + j] == code[dev[sel, 2].long() - 1 + j]).all())
        print("C5 properties ok")
    ix.free()


if "big" in what:
    # 1.5 Gbp, k=32: 1.5e9 records (the reference's int coordinates allow < 2^31), ~56 GB of HBM in flight
    Lb = 1_500_000_000
    t0 = time.time()
    sb = synth.generate(Lb, 0xB16, repeat=0.2, tandem=0.02, homo=0.001, lower=0.2, n_gaps=200, gap_max=200000, n_single=5000,
                        tandem_len_min=100, tandem_len_max=5000, homo_len_max=80)
    print(f"big: {Lb} bases generated in {time.time() - t0:.1f}s")
    db = torch.from_numpy(sb).cuda()
    ix, ms = timed(lambda: kh.make_kmer_hash(db, 32))
    ix.free()
    kh.profile(reset=True)                                 # the first launches of a process include lazy module loading
    ix, ms = timed(lambda: kh.make_kmer_hash(db, 32))
    U, N, P = ix.sizes
    print(f"big build k=32: {ms:.1f} ms, N={N} U={U} P={P} -> {N / ms / 1e6:.2f} G k-mers/s")
    show_profile("big build")
    keys = torch.empty(U, dtype=torch.int64, device="cuda")
    _lib.check(L.kmg_kmers_u64(ix._handle(), keys.data_ptr()))
    assert int(torch.unique(keys).numel()) == U                       # distinct (grouped order: not ascending)
    cnt = torch.empty(U, dtype=torch.int32, device="cuda")
    _lib.check(L.kmg_counts(ix._handle(), cnt.data_ptr()))
    assert int(cnt.sum(dtype=torch.int64)) == N and int(cnt.min()) >= 1
    del keys
    # positions in two halves (an N x 2 int32 matrix is 12 GB): every position appears exactly once
    seen = torch.zeros(Lb + 1, dtype=torch.uint8, device="cuda")
    code = ((db >> 1) & 3)
    isn = ((db | 0x20) == ord("n"))
    st = C.c_void_p()
    half = N // 2
    from kmer_hasher_b200 import kmer_pos
    pos = torch.empty((N, 2), dtype=torch.int32, device="cuda")
    _lib.check(L.kmg_positions(ix._handle(), pos.data_ptr()))
    p = pos[:, 1].long()
    seen.index_add_(0, p, torch.ones(1, dtype=torch.uint8, device="cuda").expand(N))
    assert int(seen.max()) == 1 and int(seen.sum(dtype=torch.int64)) == N
    same = pos[1:, 0] == pos[:-1, 0]
    assert bool((pos[1:, 1][same] > pos[:-1, 1][same]).all()) and bool((pos[1:, 0] >= pos[:-1, 0]).all())
    sel = torch.randint(0, N, (2_000_000,), device="cuda")
    st0 = p[sel] - 1
    for j in range(32):
        assert not bool(isn[st0 + j].any())
    grp = pos[sel, 0].long() - 1
    w = torch.zeros_like(st0)
    for j in range(32):
        w = (w << 2) | code[st0 + j].long()
    keys = torch.empty(U, dtype=torch.int64, device="cuda")
    _lib.check(L.kmg_kmers_u64(ix._handle(), keys.data_ptr()))
    assert bool((w == keys[grp]).all())
    print("big properties ok: keys distinct, counts sum to N, every position exactly once, lists ascending, sampled windows re-encode to their k-mer")
    ix.free()
