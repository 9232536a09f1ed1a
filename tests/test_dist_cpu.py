"""CPU test of the N>1 host logic (kmer_hasher_b200/dist.py) with world_size 2 and 3 over gloo.

The device operations are replaced by a test double built on the oracle, so what is checked here
is the sharding itself: shard bounds and halo exchange, splitter choice, record routing through
all_to_all, source-order concatenation (ascending positions per k-mer), global k-mer numbering, and the
routed probe.  The CUDA engine is covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, random_dna


class OracleEngine:
    """Test double for dist.CudaEngine: same interface, numpy + the plain-C oracle."""

    def __init__(self):
        from oracle import Oracle
        self.o = Oracle()
        self.device = torch.device("cpu")

    def upload(self, a):
        return torch.from_numpy(np.ascontiguousarray(a))

    @staticmethod
    def _bytes(shard):
        return shard.numpy().astype(np.uint8)

    def sample(self, shard, g0, g1, L, s0, s1, k, n):
        b = self._bytes(shard)
        nstarts = max(min(s1, L - k + 1) - s0, 0)
        out = np.zeros(n, np.uint64)
        for i in range(n):
            q = (i * nstarts) // n if nstarts > 0 else 0
            w = 0
            for j in range(k):
                o = s0 - g0 + q + j
                c = int(b[o]) if o < len(b) else 0
                w = ((w << 2) | ((c >> 1) & 3)) & ((1 << (2 * k)) - 1)
            out[i] = w
        return torch.from_numpy(out.view(np.int64))

    def partition(self, shard, g0, g1, L, s0, s1, k, splitters, nparts):
        b = self._bytes(shard)
        if g1 < L:                                   # not the global end: keep the local end-of-string rule away
            b = np.concatenate([b, np.frombuffer(b"A", np.uint8)])
        keys, pos = self.o.windows(b, k)
        gpos = pos.astype(np.int64) + g0              # local 1-based -> global 1-based
        keep = (gpos - 1 >= s0) & (gpos - 1 < s1)
        keys, gpos = keys[keep], gpos[keep]
        owner = np.searchsorted(np.asarray(splitters, np.uint64), keys, side="right") if nparts > 1 else np.zeros(len(keys), np.int64)
        order = np.argsort(owner, kind="stable")
        counts = np.bincount(owner, minlength=nparts).tolist()
        return (torch.from_numpy(keys[order].view(np.int64).copy()), torch.from_numpy(gpos[order].astype(np.int32)), counts)

    def build_records(self, keys, pos, n, k):
        ix = self.o.build_from_records(keys.numpy()[:n].view(np.uint64), pos.numpy()[:n], k)
        ix.sizes = (ix.U, ix.N, ix.P)
        ix.sizes_un = (ix.U, ix.N)
        ix.free = ix.close
        return ix

    def query_records(self, index, keys, coords, n):
        e = index.extract(2 | 8)
        uk, cnt = e["keys"], e["count"].astype(np.int64)
        start = np.concatenate([[0], np.cumsum(cnt)])
        p = e["pos"][1::2]
        rows = []
        kk = keys.numpy()[:n].view(np.uint64)
        cc = coords.numpy()[:n]
        at = np.searchsorted(uk, kk)
        for a, key, c in zip(at, kk, cc):
            if a < len(uk) and uk[a] == key:
                for j in p[start[a]:start[a + 1]]:
                    rows.append((int(c), int(j)))
        return torch.tensor(rows, dtype=torch.int32).reshape(-1, 2)


def _worker(rank, world, port, seq, query, k, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kmer_hasher_b200 import dist as kdist
        eng = OracleEngine()
        L = len(seq)
        s0, s1, _, _ = kdist.shard_bounds(L, world, rank, k)
        ix = kdist.sharded_build(seq[s0:s1], L, k, eng, n_samples=64)
        e = ix.local.extract(2 | 8)
        pos = e["pos"].reshape(-1, 2).copy()
        pos[:, 0] += ix.i_offset
        Lq = len(query)
        q0, q1, _, _ = kdist.shard_bounds(Lq, world, rank, k)
        rows = kdist.sharded_query(ix, query[q0:q1], Lq, k)
        ret[rank] = dict(keys=e["keys"], count=e["count"], pos=pos, rows=rows.numpy(), U_all=ix.U_all, N_total=ix.N_total,
                         splitters=ix.splitters)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,k,n", [(2, 12, 20000), (3, 32, 9001), (2, 5, 3000)])
def test_sharded_build_and_query_match_single_index(oracle, world, k, n):
    seq = random_dna(n, 5 + world, p_n=0.002, p_lower=0.2, n_runs=6)
    per = (n + world - 1) // world
    seq[per - 3:per + 2] = np.frombuffer(b"ACNGT", np.uint8)       # a breaker right at the first cut
    seq[n - k - 1] = ord("N")                                       # final run of exactly k: end-of-string rule
    query = random_dna(n // 2, 99, p_n=0.001)
    query[100:100 + n // 4] = seq[50:50 + n // 4]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), seq, query, k, ret), nprocs=world, join=True)
    whole = oracle.build(seq, k)
    want = whole.extract(2 | 8)
    keys = np.concatenate([ret[r]["keys"] for r in range(world)])
    assert np.array_equal(keys, want["keys"])                       # owners hold consecutive key ranges
    assert np.array_equal(np.concatenate([ret[r]["count"] for r in range(world)]), want["count"])
    assert np.array_equal(np.concatenate([ret[r]["pos"] for r in range(world)]).ravel(), want["pos"])
    assert ret[0]["N_total"] == whole.N and sum(ret[0]["U_all"]) == whole.U
    # probe: merge the owners' row blocks by i (stable) -> the reference order
    rows = np.concatenate([ret[r]["rows"] for r in range(world)])
    rows = rows[np.argsort(rows[:, 0], kind="stable")]
    assert np.array_equal(rows.ravel(), whole.query(query, k))


def test_shard_bounds_cover_every_window():
    from kmer_hasher_b200.dist import shard_bounds
    for L in (0, 1, 31, 32, 33, 1000, 12345):
        for world in (1, 2, 3, 8):
            for k in (1, 16, 32):
                starts = []
                for r in range(world):
                    s0, s1, g0, g1 = shard_bounds(L, world, r, k)
                    assert 0 <= g0 <= s0 <= s1 <= L and g1 <= L
                    assert g0 <= max(s0 - 1, 0) and g1 >= min(L, s1 + k - 1)
                    starts += list(range(s0, s1))
                assert starts == list(range(L))


def test_choose_splitters_quantiles():
    from kmer_hasher_b200.dist import choose_splitters
    s = np.arange(1000, dtype=np.uint64)[::-1].copy()
    assert choose_splitters(s, 4).tolist() == [250, 500, 750]
    assert choose_splitters(s, 1).size == 0
    big = np.array([2**63 + 5, 1, 2**64 - 1, 7], np.uint64)        # unsigned order, not int64 order
    assert choose_splitters(big, 2).tolist() == [2**63 + 5]


# ---- the region exchange (grouped sharded build) with a test double ------------------------------------------------
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def np_mix64(x):
    x = x.astype(np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33); x *= np.uint64(0xff51afd7ed558ccd)
        x ^= x >> np.uint64(33); x *= np.uint64(0xc4ceb9fe1a85ec53)
        x ^= x >> np.uint64(33)
    return x


def np_unmix64(x):
    x = x.astype(np.uint64).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(33); x *= np.uint64(0x9cb4b2f8129337db)
        x ^= x >> np.uint64(33); x *= np.uint64(0x4f74430c22a54005)
        x ^= x >> np.uint64(33)
    return x


class _Box:
    pass


class GlooExchange:
    """Stand-in for dist.PeerExchange: `barrier` performs the exchange that peer stores over NVLink do on the GPU."""

    def __init__(self, region_cap):
        self.region_cap = region_cap
        self.slot = None

    def next_slot(self):
        self.slot = _Box()
        self.slot.outbox, self.slot.inbox = None, None
        return self.slot

    def barrier(self):
        world = dist.get_world_size()
        out = self.slot.outbox                                # per owner: (mixed keys u64, payload i32)
        send_counts = torch.tensor([len(o[0]) for o in out], dtype=torch.int64)
        recv_counts = torch.empty(world, dtype=torch.int64)
        dist.all_to_all_single(recv_counts, send_counts)
        rc, sc = recv_counts.tolist(), send_counts.tolist()
        sk = torch.from_numpy(np.concatenate([o[0] for o in out]).view(np.int64).copy())
        sp = torch.from_numpy(np.concatenate([o[1] for o in out]).astype(np.int32))
        rk, rp = torch.empty(sum(rc), dtype=torch.int64), torch.empty(sum(rc), dtype=torch.int32)
        dist.all_to_all_single(rk, sk, rc, sc)
        dist.all_to_all_single(rp, sp, rc, sc)
        offs = np.concatenate([[0], np.cumsum(rc)])
        self.slot.inbox = [(rk.numpy()[offs[r]:offs[r + 1]].view(np.uint64), rp.numpy()[offs[r]:offs[r + 1]]) for r in range(world)]

    def agree(self, status, sizes=(0, 0)):
        world = dist.get_world_size()
        t = torch.tensor([int(status), int(sizes[0]), int(sizes[1])], dtype=torch.int64)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        a = torch.stack(allt).numpy()
        return int(a[:, 0].max()), a[:, 1].tolist(), a[:, 2].tolist()


class RegionEngine(OracleEngine):
    """OracleEngine + the region-exchange interface of dist.CudaEngine."""

    def shard_pack(self, own, k, n_samples, order=0):
        b = self._bytes(own)
        pack = np.zeros(48 + 8 * n_samples, np.uint8)
        pack[:min(k - 1, len(b))] = b[:k - 1]
        if len(b):
            pack[40] = b[-1]
        return torch.from_numpy(pack)

    def shard_open_packed(self, own, L, world, rank, k, n_samples, allpack, order=0, splitters=True):
        from kmer_hasher_b200.dist import shard_bounds
        packs = allpack.numpy().reshape(world, -1)
        s0, s1, g0, g1 = shard_bounds(L, world, rank, k)
        b = self._bytes(own)
        left = packs[rank - 1][40:41] if g0 < s0 else np.empty(0, np.uint8)
        need, right, r = g1 - s1, [], rank + 1
        while need > 0 and r < world:
            rs0, rs1, _, _ = shard_bounds(L, world, r, k)
            take = min(need, rs1 - rs0, k - 1)
            right.append(packs[r][:take]); need -= take; r += 1
        h = _Box()
        h.bytes, h.g0, h.g1, h.s0, h.s1, h.L, h.k = np.concatenate([left, b] + right), g0, g1, s0, s1, L, k
        return h, None

    def shard_close(self, h):
        pass

    def shard_scatter_ranges(self, h, nparts, rank, slot, region_cap, pos_add):
        b = h.bytes
        if h.g1 < h.L:
            b = np.concatenate([b, np.frombuffer(b"A", np.uint8)])
        keys, pos = self.o.windows(b, h.k)
        gpos = pos.astype(np.int64) + h.g0
        keep = (gpos - 1 >= h.s0) & (gpos - 1 < h.s1)
        mixed, gpos = np_mix64(keys[keep]), gpos[keep] + pos_add
        owner = ((mixed.astype(object) * nparts) >> 64).astype(np.int64) if len(mixed) else np.zeros(0, np.int64)   # __umul64hi(h, nparts)
        slot.outbox = [(mixed[owner == o], gpos[owner == o].astype(np.int32)) for o in range(nparts)]

    def build_regions(self, slot, region_cap, nparts, k):
        from kmer_hasher_b200._lib import KmgError
        if any(len(h) > region_cap for h, _ in slot.inbox):
            raise KmgError(-3, "an owner received more than the exchange capacity")
        keys = np_unmix64(np.concatenate([h for h, _ in slot.inbox]))
        pos = np.concatenate([p for _, p in slot.inbox])
        return self.build_records(torch.from_numpy(keys.view(np.int64).copy()), torch.from_numpy(pos), len(keys), k)

    def query_regions(self, index, slot, region_cap, nparts, count_only=False):
        keys = np_unmix64(np.concatenate([h for h, _ in slot.inbox]))
        co = np.concatenate([p for _, p in slot.inbox])
        return self.query_records(index, torch.from_numpy(keys.view(np.int64).copy()), torch.from_numpy(co), len(keys))


def _region_worker(rank, world, port, seq, query, k, region_cap, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kmer_hasher_b200 import dist as kdist
        eng, L = RegionEngine(), len(seq)
        s0, s1, _, _ = kdist.shard_bounds(L, world, rank, k)
        ix = kdist.sharded_build_ranges(seq[s0:s1], L, k, eng, GlooExchange(region_cap))
        e = ix.local.extract(2 | 8)
        pos = e["pos"].reshape(-1, 2).copy()
        pos[:, 0] += ix.i_offset                              # global k-mer numbers, owners in rank order
        rows = None
        if ix.ranges:
            Lq = len(query)
            q0, q1, _, _ = kdist.shard_bounds(Lq, world, rank, k)
            rows = kdist.sharded_query_ranges(ix, query[q0:q1], Lq, k, GlooExchange(region_cap)).numpy()
        ret[rank] = dict(keys=e["keys"], count=e["count"], pos=pos, rows=rows, ranges=ix.ranges, row_offset=ix.row_offset,
                         N_total=ix.N_total)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,k,n,cap", [(2, 32, 20000, 1 << 20), (3, 21, 9001, 1 << 20), (2, 27, 6000, 500)])
def test_region_exchange_build_and_query(oracle, world, k, n, cap):
    """sharded_build_ranges / sharded_query_ranges over gloo: halo packs, owners = equal ranges of the mixed key, per-source
    regions, collective agreement on the outcome.  cap = 500 makes every region overflow: all ranks must then take the
    general path together (no hang, same result)."""
    seq = random_dna(n, 7 + world, p_n=0.002, p_lower=0.2, n_runs=5)
    per = (n + world - 1) // world
    seq[per - 3:per + 2] = np.frombuffer(b"ACNGT", np.uint8)
    seq[n - k - 1] = ord("N")
    seq[200:200 + 3 * k] = ord("a")                                  # a repeat: many copies of one k-mer go to one owner
    query = random_dna(n // 2, 99, p_n=0.001)
    query[100:100 + n // 4] = seq[50:50 + n // 4]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_region_worker, args=(world, _free_port(), seq, query, k, cap, ret), nprocs=world, join=True)
    whole = oracle.build(seq, k)
    want = whole.extract(2 | 8)
    assert all(ret[r]["ranges"] == (cap > 500) for r in range(world))
    keys = np.concatenate([ret[r]["keys"] for r in range(world)])
    cnt = np.concatenate([ret[r]["count"] for r in range(world)])
    pos = np.concatenate([ret[r]["pos"] for r in range(world)])
    assert ret[0]["N_total"] == whole.N and [ret[r]["row_offset"] for r in range(world)] == np.concatenate([[0], np.cumsum([len(ret[r]["pos"]) for r in range(world)])])[:-1].tolist()
    order = np.argsort(keys, kind="stable")                          # owners hold ranges of the MIXED key: canonicalise
    rank = np.empty(len(keys), np.int64); rank[order] = np.arange(len(keys))
    assert np.array_equal(keys[order], want["keys"]) and np.array_equal(cnt[order], want["count"])
    new_i = rank[pos[:, 0].astype(np.int64) - 1] + 1
    o2 = np.argsort(new_i, kind="stable")
    canon = pos[o2].copy(); canon[:, 0] = new_i[o2]
    assert np.array_equal(canon.ravel(), want["pos"])
    if cap > 500:
        rows = np.concatenate([ret[r]["rows"] for r in range(world)])
        rows = rows[np.argsort(rows[:, 0], kind="stable")]
        assert np.array_equal(rows.ravel(), whole.query(query, k))
