"""Sharded k-mer position index: one process per GPU, torch.distributed for the plumbing.

The path shards with ONE exchange step:

  1. the global sequence is cut into G contiguous shards; a rank needs k-1 bytes of its right
     neighbour (windows straddling the cut) and one byte of its left neighbour (end-of-string rule),
     fetched by a halo exchange;
  2. every rank samples window keys, the samples are all-gathered and G-1 splitters are taken at
     the quantiles, so key ranges are balanced on repeat-rich input;
  3. the shard is encoded and its (key,pos) records are grouped by owner (CUDA, libkmergpu
     kmg_shard_partition: the sort pass with a key-range bin function, order-preserving);
  4. one all-to-all of the records over NVLink; receivers get the groups in source-rank order, i.e.
     ascending positions, so the stable local sort keeps every k-mer's list ascending;
  5. each owner sorts its records and builds its CSR slice (kmg_build_records); the global 1-based
     k-mer index is local rank + the exclusive prefix of U over ranks.

That is the general path (sharded_build / sharded_query).  The product path on an NVLink node is
sharded_build_p2p / sharded_query_p2p: steps 1-2 collapse into ONE small all-gather (halo bytes + a sorted
key sample per rank, splitters chosen on the device), and steps 3-4 into ONE kernel: the partitioning
pass writes every record straight into its owner's arrays, mapped from the peer GPUs with CUDA IPC
(PeerExchange), so there is no staging buffer and no all-to-all; the host never waits for the device
before the finished index is read.  If peer memory cannot be mapped on every rank, or an owner's share
exceeds the exchange capacity, all ranks take the general path together.

The device work sits behind a small engine object so the host logic can be exercised on CPU with
gloo (tests/test_dist_cpu.py plugs in a test double); the product engine is CudaEngine.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch
import torch.distributed as dist


class CudaEngine:
    """Device operations through the C ABI of libkmergpu (no CPU path)."""

    def __init__(self, device: torch.device):
        from . import _lib
        self._lib = _lib
        self.L = _lib.load()
        self.device = device
        _lib.check(self.L.kmg_set_device(device.index or 0))
        _lib.check(self.L.kmg_set_stream(torch.cuda.current_stream(device).cuda_stream))

    def upload(self, a: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(a)).to(self.device, non_blocking=False)

    def sample(self, shard, g0, g1, L, s0, s1, k, n):
        out = torch.empty(n, dtype=torch.int64, device=self.device)
        self._lib.check(self.L.kmg_shard_sample(shard.data_ptr(), g0, g1, L, s0, s1, k, n, out.data_ptr()))
        return out

    def partition(self, shard, g0, g1, L, s0, s1, k, splitters: np.ndarray, nparts: int):
        cap = max(s1 - s0, 1)
        keys = torch.empty(cap, dtype=torch.int64, device=self.device)
        pos = torch.empty(cap, dtype=torch.int32, device=self.device)
        counts = (C.c_uint64 * nparts)()
        spl = np.ascontiguousarray(splitters, np.uint64)
        self._lib.check(self.L.kmg_shard_partition(shard.data_ptr(), g0, g1, L, s0, s1, k,
                                                   spl.ctypes.data_as(C.POINTER(C.c_uint64)), nparts,
                                                   keys.data_ptr(), pos.data_ptr(), counts))
        return keys, pos, [int(c) for c in counts]

    def build_records(self, keys, pos, n, k):
        from . import KmerHash
        h = C.c_void_p()
        self._lib.check(self.L.kmg_build_records(keys.data_ptr(), pos.data_ptr(), n, k, C.byref(h)))
        return KmerHash(h.value, k)

    # ---- peer-memory path (NVLink P2P, no host synchronisation before the index is read) ----------------
    def shard_open(self, shard, g0, g1, L, s0, s1, k):
        h = C.c_void_p()
        self._lib.check(self.L.kmg_shard_open(shard.data_ptr(), g0, g1, L, s0, s1, k, C.byref(h)))
        return h

    def shard_pack(self, own, k, n_samples, order=0):
        pack = torch.empty(self.L.kmg_shard_pack_bytes(n_samples), dtype=torch.uint8, device=self.device)
        self._lib.check(self.L.kmg_shard_pack(own.data_ptr(), own.numel(), k, n_samples, order, pack.data_ptr()))
        return pack

    def shard_open_packed(self, own, L, world, rank, k, n_samples, allpack, order=0, splitters=True):
        h = C.c_void_p()
        spl = torch.empty(max(world - 1, 1), dtype=torch.int64, device=self.device) if splitters else None
        self._lib.check(self.L.kmg_shard_open_packed(own.data_ptr(), own.numel(), L, world, rank, k, n_samples, order,
                                                     allpack.data_ptr(), C.byref(h), spl.data_ptr() if splitters else None))
        return h, (spl[:world - 1] if splitters else None)

    def shard_close(self, h):
        self.L.kmg_shard_close(h)

    def sample_keys(self, h, n):
        out = torch.empty(n, dtype=torch.int64, device=self.device)
        self._lib.check(self.L.kmg_shard_sample_keys(h, n, out.data_ptr()))
        return out

    def shard_count(self, h, splitters_dev, nparts):
        out = torch.empty(nparts, dtype=torch.int64, device=self.device)
        self._lib.check(self.L.kmg_shard_count(h, splitters_dev.data_ptr(), nparts, out.data_ptr()))
        return out

    def shard_scatter(self, h, splitters_dev, nparts, rank, slot, capacity, matrix, pos_add):
        info = torch.empty(2, dtype=torch.int64, device=self.device)
        self._lib.check(self.L.kmg_shard_scatter(h, splitters_dev.data_ptr(), nparts, rank, slot.peer_keys, slot.peer_pos,
                                                 capacity, matrix.data_ptr(), pos_add, info.data_ptr()))
        return info

    def build_received(self, slot, capacity, info, k, order=1):
        from . import KmerHash
        h = C.c_void_p()
        self._lib.check(self.L.kmg_build_received(slot.keys, slot.pos, capacity, info.data_ptr(), k, order, C.byref(h)))
        return KmerHash(h.value, k)

    def query_received(self, index, slot, capacity, info, mixed=False, count_only=False):
        st, M = C.c_void_p(), C.c_uint64()
        self._lib.check(self.L.kmg_query_received(index._handle(), slot.keys, slot.pos, capacity, info.data_ptr(),
                                                  int(mixed), C.byref(st), C.byref(M)))
        if count_only:                                       # lookups + compaction + row count, rows not emitted
            self.L.kmg_query_free(st)
            return int(M.value)
        rows = torch.empty((M.value, 2), dtype=torch.int32, device=self.device)
        try:
            self._lib.check(self.L.kmg_query_emit(st, rows.data_ptr()))
        finally:
            self.L.kmg_query_free(st)
        return rows

    # ---- region exchange: owners = equal ranges of the mixed key, nothing counted or exchanged before the scatter ----
    def shard_scatter_ranges(self, h, nparts, rank, slot, region_cap, pos_add):
        self._lib.check(self.L.kmg_shard_scatter_ranges(h, nparts, rank, slot.peer_keys, slot.peer_pos, slot.peer_counts,
                                                        region_cap, pos_add))

    def build_regions(self, slot, region_cap, nparts, k):
        from . import KmerHash
        h = C.c_void_p()
        self._lib.check(self.L.kmg_build_regions(slot.keys, slot.pos, region_cap, nparts, slot.counts, k, C.byref(h)))
        return KmerHash(h.value, k)

    def query_regions(self, index, slot, region_cap, nparts, count_only=False):
        st, M = C.c_void_p(), C.c_uint64()
        self._lib.check(self.L.kmg_query_regions(index._handle(), slot.keys, slot.pos, region_cap, nparts, slot.counts,
                                                 C.byref(st), C.byref(M)))
        if count_only:
            self.L.kmg_query_free(st)
            return int(M.value)
        rows = torch.empty((M.value, 2), dtype=torch.int32, device=self.device)
        try:
            self._lib.check(self.L.kmg_query_emit(st, rows.data_ptr()))
        finally:
            self.L.kmg_query_free(st)
        return rows

    def positions_base(self, index, i_base, out):
        self._lib.check(self.L.kmg_positions_base(index._handle(), i_base, out.data_ptr()))

    def query_records(self, index, keys, coords, n):
        st, M = C.c_void_p(), C.c_uint64()
        self._lib.check(self.L.kmg_query_records(index._handle(), keys.data_ptr(), coords.data_ptr(), n,
                                                 C.byref(st), C.byref(M)))
        rows = torch.empty((M.value, 2), dtype=torch.int32, device=self.device)
        try:
            self._lib.check(self.L.kmg_query_emit(st, rows.data_ptr()))
        finally:
            self.L.kmg_query_free(st)
        return rows


_MARKS = None     # tuning runs: set to a list to collect (phase, cuda event, host time) marks of the sharded builds


def _mark(name):
    if _MARKS is not None:
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        _MARKS.append((name, e, time.perf_counter()))


class PeerUnavailable(RuntimeError):
    """Raised by PeerExchange on every rank when peer memory cannot be used; take the all-to-all path."""


class _Slot:
    """One set of receive arrays: this rank's (keys, pos) and every rank's, as the scatter sees them."""
    __slots__ = ("keys", "pos", "counts", "peer_keys", "peer_pos", "peer_counts")


class PeerExchange:
    """Receive arrays for the fused partition + exchange, mapped into every rank of the node (CUDA IPC
    over NVLink).  Two alternating slots: a peer may already scatter the next build's records while this
    rank is still sorting the current one."""

    def __init__(self, engine: "CudaEngine", capacity: int, group=None):
        """Collective.  Raises PeerUnavailable on EVERY rank if any rank cannot allocate or map the arrays
        (e.g. devices hidden from one another), so callers can fall back to the all-to-all path together."""
        self.engine, self.capacity, self.group = engine, int(capacity), group
        L, lib = engine.L, engine._lib
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self._own, self._opened = [], []
        # region exchange: every source rank has its own region of region_cap slots in every owner's arrays
        self.region_cap = self.capacity // max(world, 1)
        NB = 6                                               # per slot: keys, pos, received counts
        sizes = [self.capacity * 8, self.capacity * 4, 16 * 8] * 2
        handles = torch.zeros(NB * 64, dtype=torch.uint8)
        ok = True
        try:
            for i in range(NB):                              # slot0 keys, pos, counts, slot1 keys, pos, counts
                p, hbuf = C.c_void_p(), (C.c_ubyte * 64)()
                lib.check(L.kmg_ipc_alloc(sizes[i], C.byref(p), hbuf))
                self._own.append(p.value)
                handles[i * 64:(i + 1) * 64] = torch.frombuffer(bytearray(hbuf), dtype=torch.uint8)
        except lib.KmgError:
            ok = False
        mine = handles.to(engine.device)
        allh = torch.empty(world * NB * 64, dtype=torch.uint8, device=engine.device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        allh = allh.cpu().numpy().reshape(world, NB, 64)
        ptrs = [[0] * NB for _ in range(world)]
        if ok:
            try:
                for r in range(world):
                    for i in range(NB):
                        if r == rank:
                            ptrs[r][i] = self._own[i]
                        else:
                            p = C.c_void_p()
                            hb = (C.c_ubyte * 64).from_buffer_copy(allh[r, i].tobytes())
                            lib.check(L.kmg_ipc_open(hb, C.byref(p)))
                            self._opened.append(p.value)
                            ptrs[r][i] = p.value
            except lib.KmgError:
                ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=engine.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            for p in self._opened:
                L.kmg_ipc_close(p)
            dist.barrier(group=group)
            for p in self._own:
                L.kmg_ipc_free(p)
            self._opened, self._own = [], []
            raise PeerUnavailable("peer memory could not be mapped on every rank")
        self.slots = []
        for sidx in range(2):
            sl = _Slot()
            sl.keys, sl.pos, sl.counts = self._own[3 * sidx], self._own[3 * sidx + 1], self._own[3 * sidx + 2]
            sl.peer_keys = (C.c_void_p * world)(*[ptrs[r][3 * sidx] for r in range(world)])
            sl.peer_pos = (C.c_void_p * world)(*[ptrs[r][3 * sidx + 1] for r in range(world)])
            sl.peer_counts = (C.c_void_p * world)(*[ptrs[r][3 * sidx + 2] for r in range(world)])
            self.slots.append(sl)
        self._turn = 0
        self._token = torch.zeros(1, dtype=torch.int32, device=engine.device)

    def next_slot(self) -> _Slot:
        self._turn ^= 1
        return self.slots[self._turn]

    def barrier(self):
        """Device-side: every rank's scatter has finished before any rank reads what it received."""
        dist.all_reduce(self._token, group=self.group)

    def agree(self, status: int, sizes=(0, 0)):
        """The worst status any rank reports, and every rank's (U, N) -- ONE small all-gather: what to do next is decided
        together, never by one rank alone, and the global k-mer numbering needs no further collective."""
        world = dist.get_world_size(self.group)
        t = torch.tensor([int(status), int(sizes[0]), int(sizes[1])], dtype=torch.int64, device=self.engine.device)
        allt = torch.empty(3 * world, dtype=torch.int64, device=self.engine.device)
        dist.all_gather_into_tensor(allt, t, group=self.group)
        a = allt.cpu().numpy().reshape(world, 3)
        return int(a[:, 0].max()), a[:, 1].tolist(), a[:, 2].tolist()

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for p in self._opened:
            self.engine.L.kmg_ipc_close(p)
        dist.barrier(group=self.group)
        for p in self._own:
            self.engine.L.kmg_ipc_free(p)
        self._opened, self._own = [], []


def shard_bounds(L: int, world: int, rank: int, k: int):
    """Window starts [s0,s1) owned by `rank` and the bytes [g0,g1) it must hold."""
    per = (L + world - 1) // world
    s0 = min(rank * per, L)
    s1 = min((rank + 1) * per, L)
    g0 = max(s0 - 1, 0)
    g1 = min(L, s1 + k - 1)
    return s0, s1, g0, g1


def exchange_halo(own: torch.Tensor, L: int, k: int, rank: int, world: int, group=None) -> tuple[torch.Tensor, int, int]:
    """own = this rank's bytes [s0,s1) of the global sequence (device of the backend).  Returns the
    bytes [g0,g1) after fetching k-1 bytes from the right neighbour(s) and one byte from the left."""
    s0, s1, g0, g1 = shard_bounds(L, world, rank, k)
    per = (L + world - 1) // world
    assert own.numel() == s1 - s0
    if world == 1:
        return own, g0, g1
    # left context: last byte of the left neighbour; right context: first k-1 bytes to the right.
    # (k-1 <= 31 bytes always come from one neighbour unless shards are shorter than k; handle the
    # general case by all-gathering fixed-size heads/tails, which is tiny.)
    ht = torch.zeros(k + 1, dtype=torch.uint8, device=own.device)     # [0,k): head, [k]: last byte
    n_head = min(k - 1, own.numel())
    if n_head:
        ht[:n_head] = own[:n_head]
    if own.numel():
        ht[k] = own[-1]
    allht = torch.empty(world * (k + 1), dtype=torch.uint8, device=own.device)
    dist.all_gather_into_tensor(allht, ht, group=group)
    allht = allht.view(world, k + 1)
    heads = [allht[r, :k] for r in range(world)]
    tails = [allht[r, k:] for r in range(world)]
    tail = ht[k:]
    parts = []
    if g0 < s0:
        parts.append(tails[rank - 1] if per >= 1 else tail)
    parts.append(own)
    need = g1 - s1
    r = rank + 1
    while need > 0 and r < world:
        rs0, rs1, _, _ = shard_bounds(L, world, r, k)
        take = min(need, rs1 - rs0, k - 1)
        parts.append(heads[r][:take])
        need -= take
        r += 1
    return torch.cat(parts), g0, g1


def choose_splitters(samples_all: np.ndarray, world: int) -> np.ndarray:
    """G-1 splitters at the quantiles of the gathered sample (uint64, ascending)."""
    s = np.sort(samples_all.astype(np.uint64))
    if world == 1 or s.size == 0:
        return np.zeros(0, np.uint64)
    idx = [(i * s.size) // world for i in range(1, world)]
    return s[idx].astype(np.uint64)


def device_splitters(samples_all: torch.Tensor, world: int) -> torch.Tensor:
    """choose_splitters on the device (keys are uint64 held in int64 tensors: flip the sign bit to sort)."""
    flip = torch.tensor(-2**63, dtype=torch.int64, device=samples_all.device)
    s, _ = torch.sort(samples_all ^ flip)
    idx = torch.tensor([(i * s.numel()) // world for i in range(1, world)], dtype=torch.int64, device=s.device)
    return (s[idx] ^ flip).contiguous()


class ShardedIndex:
    """One rank's slice of a sharded index: the k-mers whose key falls in this rank's range.
    U_all / N_all (every rank's sizes, for the global 1-based k-mer index) are gathered on first use:
    a collective, so all ranks must ask together."""

    def __init__(self, local, k, rank, world, U_all, N_all, splitters, engine, splitters_dev=None, group=None):
        self.local, self.k, self.rank, self.world = local, k, rank, world
        self._U_all, self._N_all, self._splitters, self.engine = U_all, N_all, splitters, engine
        self.splitters_dev, self.group = splitters_dev, group
        self.mixed, self.order = False, 1                    # set by the peer-memory build when owners hold ranges of the mix
        self.ranges = False                                  # owners are equal ranges of the mixed key (region exchange)

    def _gather_sizes(self):
        if self._U_all is None:
            U, N = self.local.sizes_un
            dev = self.engine.device
            sizes = torch.tensor([U, N], dtype=torch.int64, device=dev)
            allsz = torch.empty(self.world * 2, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allsz, sizes, group=self.group)
            allsz = allsz.cpu().numpy().reshape(self.world, 2)
            self._U_all, self._N_all = allsz[:, 0].tolist(), allsz[:, 1].tolist()

    @property
    def U_all(self):
        self._gather_sizes()
        return self._U_all

    @property
    def N_all(self):
        self._gather_sizes()
        return self._N_all

    @property
    def i_offset(self):                                   # global k-mer index = local + i_offset
        return int(sum(self.U_all[:self.rank]))

    @property
    def row_offset(self):                                 # first row of this owner's slice of the global pos matrix
        return int(sum(self.N_all[:self.rank]))

    def kmer_pos(self, flag: int = 2 | 8, out: dict | None = None) -> dict:
        """kmer.pos of the sharded index, this owner's slice (SURVEY.md 8e "Extraction"): the k-mer number i is GLOBAL
        (owners in rank order: i = local number + distinct k-mers of the owners before), so the owners' slices written at
        `row_offset` / `i_offset` into one matrix are the index's kmer.pos.  Collective on first use (sizes are gathered).
        out: optional preallocated {"pos": (N_local, 2) int32, "count": (U_local,) int32} (device or host)."""
        import kmer_hasher_b200 as kh
        U, N = self.local.sizes_un
        i_base = self.i_offset
        out = out or {}
        res = {"pos": None, "count": None, "i_offset": i_base, "row_offset": self.row_offset}
        if flag & 2:
            a = out.get("pos")
            if a is None:
                a = torch.empty((max(N, 1), 2), dtype=torch.int32, device=self.engine.device)
            ptr = a.ctypes.data if isinstance(a, np.ndarray) else a.data_ptr()
            self.engine._lib.check(self.engine.L.kmg_positions_base(self.local._handle(), i_base, ptr))
            res["pos"] = a[:N]
        if flag & 8:
            res["count"] = kh.kmer_pos(self.local, 8, out={"count": out["count"]} if "count" in out else None)["count"]
        return res

    @property
    def U_total(self):
        return int(sum(self.U_all))

    @property
    def N_total(self):
        return int(sum(self.N_all))

    @property
    def splitters(self) -> np.ndarray:
        if self._splitters is None:                        # the peer-memory build keeps them on the device
            self._splitters = self.splitters_dev.cpu().numpy().view(np.uint64)
        return self._splitters

    def free(self):
        self.local.free()


def sharded_build(own_bytes, L: int, k: int, engine, group=None, n_samples: int = 4096) -> ShardedIndex:
    """Collective: every rank passes its bytes [s0,s1) of the global sequence (numpy uint8 or a tensor
    on the engine's device).  Returns this rank's ShardedIndex."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    s0, s1, _, _ = shard_bounds(L, world, rank, k)
    own = own_bytes if isinstance(own_bytes, torch.Tensor) else engine.upload(np.asarray(own_bytes, np.uint8))
    shard, g0, g1 = exchange_halo(own, L, k, rank, world, group)

    # splitters from an all-gathered key sample
    if world > 1:
        smp = engine.sample(shard, g0, g1, L, s0, s1, k, n_samples)
        gathered = [torch.empty_like(smp) for _ in range(world)]
        dist.all_gather(gathered, smp, group=group)
        allsmp = torch.cat(gathered).cpu().numpy().view(np.uint64)
        splitters = choose_splitters(allsmp, world)
    else:
        splitters = np.zeros(0, np.uint64)

    keys, pos, counts = engine.partition(shard, g0, g1, L, s0, s1, k, splitters, world)
    n_local = sum(counts)
    if world == 1:
        local = engine.build_records(keys, pos, n_local, k)
        U, N = local.sizes_un
        return ShardedIndex(local, k, 0, 1, [U], [N], splitters, engine)

    # counts all-gather sizes the receive buffers; then ONE all-to-all per array
    send_counts = torch.tensor(counts, dtype=torch.int64, device=keys.device)
    all_counts = [torch.empty_like(send_counts) for _ in range(world)]
    dist.all_gather(all_counts, send_counts, group=group)
    matrix = torch.stack(all_counts).cpu().numpy()            # [source][owner]
    recv_counts = [int(matrix[src][rank]) for src in range(world)]
    n_recv = sum(recv_counts)
    rkeys = torch.empty(max(n_recv, 1), dtype=torch.int64, device=keys.device)
    rpos = torch.empty(max(n_recv, 1), dtype=torch.int32, device=keys.device)
    dist.all_to_all_single(rkeys[:n_recv], keys[:n_local], recv_counts, counts, group=group)
    dist.all_to_all_single(rpos[:n_recv], pos[:n_local], recv_counts, counts, group=group)
    del keys, pos
    local = engine.build_records(rkeys, rpos, n_recv, k)
    U, N = local.sizes_un
    sizes = torch.tensor([U, N], dtype=torch.int64, device=rkeys.device)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    return ShardedIndex(local, k, rank, world, all_sizes[:, 0].tolist(), all_sizes[:, 1].tolist(), splitters, engine)


def sharded_build_p2p(own_bytes, L: int, k: int, engine: CudaEngine, xchg: PeerExchange, group=None,
                      n_samples: int = 2048, order: int = 0) -> ShardedIndex:
    """sharded_build with the exchange fused into the partitioning pass: records go straight into the
    owners' arrays over NVLink (kmg_shard_scatter); the host never waits between the halo and the
    finished index.  Falls back to sharded_build if an owner's share exceeds the exchange capacity."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)

    def ev(name):
        if _MARKS is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            _MARKS.append((name, e, time.perf_counter()))
    ev("start")
    own = own_bytes if isinstance(own_bytes, torch.Tensor) else engine.upload(np.asarray(own_bytes, np.uint8))
    # order 0 (KMG_ORDER_GROUPED, as make.kmer.hash defaults to): for k >= 25 the sample, the owner ranges and the
    # records are those of the mixed key and every owner builds a grouped index; order 1: ascending keys
    pack = engine.shard_pack(own, k, n_samples, order)       # halo bytes + sorted splitter sample: ONE exchange
    allpack = torch.empty(world * pack.numel(), dtype=torch.uint8, device=engine.device)
    dist.all_gather_into_tensor(allpack, pack, group=group)
    ev("halo")
    sh, spl = engine.shard_open_packed(own, L, world, rank, k, n_samples, allpack, order)
    try:
        ev("splitters")
        counts = engine.shard_count(sh, spl, world)
        matrix = torch.empty(world * world, dtype=torch.int64, device=engine.device)
        dist.all_gather_into_tensor(matrix, counts, group=group)
        ev("counts")
        slot = xchg.next_slot()
        info = engine.shard_scatter(sh, spl, world, rank, slot, xchg.capacity, matrix, 0)
        ev("scatter")
        xchg.barrier()
        ev("barrier")
        from ._lib import KmgError
        status = 0
        try:
            local = engine.build_received(slot, xchg.capacity, info, k, order)
        except KmgError as e:
            if e.code not in (-3, -7):
                raise
            local, status = None, -e.code
    finally:
        engine.shard_close(sh)
    ev("built")
    # -3: an owner overflowed or its fix-up lists did (the latter is local to one rank); -7: the position-order check
    # failed on one device.  The ranks agree on the worst status and act together: a rank must never enter the general
    # path's collectives alone.
    worst, U_all, N_all = xchg.agree(status, local.sizes_un if local is not None else (0, 0))
    if worst:
        if local is not None:
            local.free()
        if worst == 7:                                      # that device now ranks by bitmap match: same path again
            return sharded_build_p2p(own, L, k, engine, xchg, group, n_samples, order)
        return sharded_build(own, L, k, engine, group)
    ix = ShardedIndex(local, k, rank, world, U_all, N_all, None, engine, splitters_dev=spl, group=group)
    ix.mixed = engine.L.kmg_index_order(local._handle()) == 0   # owner ranges are those of the mixed key
    ix.order = order
    return ix


def ranges_supported(k: int, order: int = 0) -> bool:
    """Region exchange carries mixed keys: the grouped build's domain (make.kmer.hash default order, k >= 21)."""
    return order == 0 and k >= 21


def sharded_build_ranges(own_bytes, L: int, k: int, engine, xchg, group=None) -> ShardedIndex:
    """The sharded build for the grouped order (k >= 21): owners are `world` EQUAL ranges of the mixed key, which is uniform
    whatever the sequence's composition, so nothing is sampled, counted or exchanged before the records move.  Per build:
    ONE small all-gather (halo bytes), the fused encode + partition + scatter over NVLink into this rank's own region of
    every owner's arrays (kmg_shard_scatter_ranges; its last tile leaves the per-owner counts with the owners), a
    one-word all-reduce as the barrier, the owner build straight from the regions (kmg_build_regions), and one status
    all-reduce so that every rank takes the same path if any owner's region overflowed."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    _mark("start")
    own = own_bytes if isinstance(own_bytes, torch.Tensor) else engine.upload(np.asarray(own_bytes, np.uint8))
    pack = engine.shard_pack(own, k, 2, 0)                     # halo bytes (the two sample keys are not used)
    allpack = torch.empty(world * pack.numel(), dtype=torch.uint8, device=engine.device)
    dist.all_gather_into_tensor(allpack, pack, group=group)
    _mark("halo")
    sh, _ = engine.shard_open_packed(own, L, world, rank, k, 2, allpack, 0, splitters=False)
    from ._lib import KmgError
    status, local = 0, None
    try:
        slot = xchg.next_slot()
        engine.shard_scatter_ranges(sh, world, rank, slot, xchg.region_cap, 0)
        _mark("scatter")
        xchg.barrier()
        _mark("barrier")
        try:
            local = engine.build_regions(slot, xchg.region_cap, world, k)
        except KmgError as e:
            if e.code not in (-3, -7):
                raise
            status = -e.code
    finally:
        engine.shard_close(sh)
    _mark("built")
    worst, U_all, N_all = xchg.agree(status, local.sizes_un if local is not None else (0, 0))
    _mark("agreed")
    if worst:
        if local is not None:
            local.free()
        if worst == 7:                                         # a device switched to the bitmap rank variant: once more
            return sharded_build_ranges(own, L, k, engine, xchg, group)
        return sharded_build(own, L, k, engine, group)         # a region overflowed (a huge repeat): the exact-size path
    ix = ShardedIndex(local, k, rank, world, U_all, N_all, None, engine, group=group)
    ix.mixed, ix.ranges, ix.order = True, True, 0
    return ix


def sharded_query_ranges(index: ShardedIndex, own_query_bytes, Lq: int, k: int, xchg, group=None, count_only: bool = False):
    """seq.kmer.pos against an index built by sharded_build_ranges: query (mixed key, i) records go to the key's owner
    through the same region scatter; each owner returns its (i,j) rows ordered by i then j."""
    engine = index.engine
    world, rank = index.world, index.rank
    own = own_query_bytes if isinstance(own_query_bytes, torch.Tensor) else engine.upload(np.asarray(own_query_bytes, np.uint8))
    pack = engine.shard_pack(own, k, 2, 0)
    allpack = torch.empty(world * pack.numel(), dtype=torch.uint8, device=engine.device)
    dist.all_gather_into_tensor(allpack, pack, group=group)
    sh, _ = engine.shard_open_packed(own, Lq, world, rank, k, 2, allpack, 0, splitters=False)
    try:
        slot = xchg.next_slot()
        engine.shard_scatter_ranges(sh, world, rank, slot, xchg.region_cap, k - 1)     # 1-based END (src/kmer_pos.c:127)
        xchg.barrier()
        return engine.query_regions(index.local, slot, xchg.region_cap, world, count_only)
    finally:
        engine.shard_close(sh)


def sharded_query_p2p(index: ShardedIndex, own_query_bytes, Lq: int, k: int, xchg: PeerExchange, group=None,
                      count_only: bool = False):
    """sharded_query with query (key, i) records scattered straight to the key's owner over NVLink.
    count_only: return this owner's number of result rows instead of the rows (what kmg_query_begin gives)."""
    engine = index.engine
    world, rank = index.world, index.rank
    own = own_query_bytes if isinstance(own_query_bytes, torch.Tensor) else engine.upload(np.asarray(own_query_bytes, np.uint8))
    spl = index.splitters_dev
    if spl is None:
        spl = torch.from_numpy(np.ascontiguousarray(index.splitters).view(np.int64)).to(engine.device)
    # halo through the same one-exchange pack as the build (its splitter sample is not needed: two keys, ignored)
    pack = engine.shard_pack(own, k, 2, 1)
    allpack = torch.empty(world * pack.numel(), dtype=torch.uint8, device=engine.device)
    dist.all_gather_into_tensor(allpack, pack, group=group)
    sh, _ = engine.shard_open_packed(own, Lq, world, rank, k, 2, allpack, 1)
    try:
        if index.mixed:                                      # owners hold ranges of the mixed key: route by it
            engine._lib.check(engine.L.kmg_shard_set_mixed(sh, 1))
        counts = engine.shard_count(sh, spl, world)
        matrix = torch.empty(world * world, dtype=torch.int64, device=engine.device)
        dist.all_gather_into_tensor(matrix, counts, group=group)
        slot = xchg.next_slot()
        info = engine.shard_scatter(sh, spl, world, rank, slot, xchg.capacity, matrix, k - 1)   # 1-based END (src/kmer_pos.c:127)
        xchg.barrier()
        return engine.query_received(index.local, slot, xchg.capacity, info, index.mixed, count_only)
    finally:
        engine.shard_close(sh)


def sharded_query(index: ShardedIndex, own_query_bytes, Lq: int, k: int, group=None) -> torch.Tensor:
    """Collective seq.kmer.pos against a sharded index: every rank passes its slice of the query.
    Query windows are routed to the key's owner exactly like index records; each owner returns its
    (i,j) rows ordered by i then j (rows of one i live on one owner, so merging the owners' outputs by i
    reproduces the reference order)."""
    engine = index.engine
    world, rank = index.world, index.rank
    if index.mixed:
        raise ValueError("this index was built grouped (owners hold ranges of the mixed key): probe it with sharded_query_p2p")
    s0, s1, _, _ = shard_bounds(Lq, world, rank, k)
    own = own_query_bytes if isinstance(own_query_bytes, torch.Tensor) else engine.upload(np.asarray(own_query_bytes, np.uint8))
    shard, g0, g1 = exchange_halo(own, Lq, k, rank, world, group)
    keys, pos, counts = engine.partition(shard, g0, g1, Lq, s0, s1, k, index.splitters, world)
    n_local = sum(counts)
    coords = pos[:n_local] + (k - 1)                       # 1-based start -> 1-based END (src/kmer_pos.c:127)
    if world == 1:
        return engine.query_records(index.local, keys, coords, n_local)
    send_counts = torch.tensor(counts, dtype=torch.int64, device=keys.device)
    all_counts = [torch.empty_like(send_counts) for _ in range(world)]
    dist.all_gather(all_counts, send_counts, group=group)
    matrix = torch.stack(all_counts).cpu().numpy()
    recv_counts = [int(matrix[src][rank]) for src in range(world)]
    n_recv = sum(recv_counts)
    rkeys = torch.empty(max(n_recv, 1), dtype=torch.int64, device=keys.device)
    rco = torch.empty(max(n_recv, 1), dtype=torch.int32, device=keys.device)
    dist.all_to_all_single(rkeys[:n_recv], keys[:n_local], recv_counts, counts, group=group)
    dist.all_to_all_single(rco[:n_recv], coords.contiguous(), recv_counts, counts, group=group)
    return engine.query_records(index.local, rkeys, rco, n_recv)


# ---------------------------------------------------------------------------------------------------------
# bench.py's N>1 leg
# ---------------------------------------------------------------------------------------------------------
MASK64 = (1 << 64) - 1


def _allreduce_u64(vals, dev):
    """Sum of 64-bit values over the ranks, mod 2^64 (halves summed separately: no reliance on overflow behaviour)."""
    halves = []
    for v in vals:
        halves += [v & 0xFFFFFFFF, (v >> 32) & 0xFFFFFFFF]
    t = torch.tensor(halves, dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    h = t.cpu().tolist()
    return [((h[2 * i + 1] << 32) + h[2 * i]) & MASK64 for i in range(len(vals))]


def sharded_parity_sums(ix: ShardedIndex):
    """Order-independent sums of a sharded index over all owners (a collective): U, N, sum of keys, sum key * count,
    sum key * pos (mod 2^64) -- the numbers tests/golden/fullsize.json holds for the reference's index of the same
    sequence ('keys'[1], 'bind') -- plus whether every owner's position lists ascend."""
    import kmer_hasher_b200 as kh
    dev = ix.engine.device
    U, N = ix.local.sizes_un
    keys = torch.empty(max(U, 1), dtype=torch.int64, device=dev)
    ix.engine._lib.check(ix.engine.L.kmg_kmers_u64(ix.local._handle(), keys.data_ptr()))
    keys = keys[:U]
    cnt = torch.empty(max(U, 1), dtype=torch.int32, device=dev)
    pos = torch.empty((max(N, 1), 2), dtype=torch.int32, device=dev)
    kh.kmer_pos(ix.local, 2 | 8, out={"pos": pos, "count": cnt})
    cnt, pos = cnt[:U].to(torch.int64), pos[:N]
    ksum = int(keys.sum().item()) & MASK64
    b0 = int((keys * cnt).sum().item()) & MASK64
    b1 = 0
    ordered = True
    step = 1 << 26
    for a in range(0, N, step):
        rows = pos[a:a + step + 1]
        b1 = (b1 + int((keys[rows[:step, 0].to(torch.int64) - 1] * rows[:step, 1].to(torch.int64)).sum().item())) & MASK64
        same = rows[1:, 0] == rows[:-1, 0]
        ordered &= bool((rows[1:, 1][same] > rows[:-1, 1][same]).all()) and bool((rows[1:, 0] >= rows[:-1, 0]).all())
    tot = _allreduce_u64([U, N, ksum, b0, b1, 0 if ordered else 1], dev)
    return {"U": tot[0], "N": tot[1], "keys_sum": tot[2], "bind": [tot[3], tot[4]], "lists_ascending": tot[5] == 0}


def bench_sharded(args, w, k, L, steps, warm, hbm_peak, peak_src, barrier):
    import json
    import os
    import kmer_hasher_b200 as kh
    from . import synth
    from bench import ClockSampler, config_of, METRIC
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    from .affinity import bind_near_gpu
    numa = {"bound": False, "why": "KMG_NO_NUMA_BIND set"} if os.environ.get("KMG_NO_NUMA_BIND") else bind_near_gpu(dev.index or 0)
    engine = CudaEngine(dev)
    for key in ("scatter_shape", "scatter_bitmap", "sort_dbg"):          # tuning runs: KMG_TUNE_scatter_shape=1 ...
        if os.environ.get("KMG_TUNE_" + key):
            engine._lib.check(engine.L.kmg_tune(key.encode(), int(os.environ["KMG_TUNE_" + key])))
    strong = args.scaling == "strong"
    Ltot = L if strong else L * world
    if Ltot > 2**31 - 2:
        raise SystemExit("global sequence exceeds the reference's int coordinates")
    s0, s1, _, _ = shard_bounds(Ltot, world, rank, k)
    n_own = s1 - s0
    own_pin = kh.pinned_empty(max(n_own, 1), np.uint8)
    seq_all = None
    if strong:                                              # ONE sequence, cut `world` ways: every rank generates it and keeps its slice
        seq_all = synth.config_c2(L) if w["gen"] == "c2" else synth.config_c3(L)
        own_pin[:n_own] = seq_all[s0:s1]
    else:                                                   # weak: every rank draws its own L bases of an (N x L)-base sequence
        (synth.config_c2 if w["gen"] == "c2" else synth.config_c3)(n_own, out=own_pin, seed=(0xC2 if w["gen"] == "c2" else 0xC3) + 7919 * rank)
    own_host = torch.from_numpy(own_pin)[:n_own]            # page-locked by kmg_host_alloc: async DMA source
    own_dev = own_host.to(dev)

    per = (Ltot + world - 1) // world
    region_cap = int(per / world * 1.3) + 65536
    try:
        xchg = PeerExchange(engine, region_cap * world)    # receive arrays mapped into every rank (NVLink P2P)
    except PeerUnavailable:
        xchg = None                                        # every rank lands here together: NCCL all-to-all path
    use_ranges = xchg is not None and ranges_supported(k)

    def build(own):
        if use_ranges:
            return sharded_build_ranges(own, Ltot, k, engine, xchg)
        return sharded_build_p2p(own, Ltot, k, engine, xchg) if xchg is not None else sharded_build(own, Ltot, k, engine)
    ix = build(own_dev)
    U, N = ix.local.sizes_un
    ntot = ix.N_total
    used_ranges = bool(ix.ranges)
    # ---- parity of the sharded index against the reference's digests of the same sequence (strong scaling only) ----
    parity = None
    gpath = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fullsize.json")
    if strong and os.path.exists(gpath):
        gold = json.load(open(gpath)).get(w.get("golden", ""), None)
        if gold and gold.get("bases") == L and "bind" in gold:
            got = sharded_parity_sums(ix)
            parity = {"ok": bool(got["U"] == gold["U"] and got["N"] == gold["N"] and got["keys_sum"] == gold["keys"][1]
                                 and got["bind"] == list(gold["bind"]) and got["lists_ascending"]),
                      "what": "U, N, sum of keys, sum key*count, sum key*pos (mod 2^64) over all owners == the reference engine's "
                              "(tests/golden/fullsize.json, made by tests/golden/make_fullsize.py); every owner's position lists ascend",
                      "got": got, "want": {"U": gold["U"], "N": gold["N"], "keys_sum": gold["keys"][1], "bind": list(gold["bind"])}}
    ix.free()
    pos_dev = torch.empty((max(N, 1), 2), dtype=torch.int32, device=dev)
    cnt_dev = torch.empty(max(U, 1), dtype=torch.int32, device=dev)
    pos_pin, cnt_pin = kh.pinned_empty((max(N, 1), 2), np.int32), kh.pinned_empty(max(U, 1), np.int32)

    def step_device():
        ix = build(own_dev)
        ix.kmer_pos(2 | 8, out={"pos": pos_dev, "count": cnt_dev})      # this owner's slice, global k-mer numbers
        _mark("kmer.pos")
        ix.free()
        _mark("freed")

    def step_e2e():
        ix = build(own_host.to(dev, non_blocking=True))
        ix.kmer_pos(2 | 8, out={"pos": pos_pin, "count": cnt_pin})
        ix.free()

    def timed(fn, n):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)           # max over ranks
        return float(t.item())

    for _ in range(warm):
        step_device()
    kh.profile(enable=True, reset=True)
    kh.profile(reset=True)
    l0 = kh.launch_count()
    with ClockSampler(dev.index or 0) as clk:
        ms = timed(step_device, steps)
    launches = kh.launch_count() - l0
    prof = kh.profile(enable=False)
    kh.profile(reset=True)
    # where a step's time goes on rank 0 (device time between marks, host time to enqueue): three more steps, untimed
    global _MARKS
    phases = None
    if use_ranges:
        _MARKS = []
        acc = {}
        for _ in range(3):
            del _MARKS[:]
            barrier()
            step_device()
            torch.cuda.synchronize()
            for (n0, e0, h0), (n1, e1, h1) in zip(_MARKS[:-1], _MARKS[1:]):
                a = acc.setdefault(n1, [0.0, 0.0])
                a[0] += e0.elapsed_time(e1) / 3
                a[1] += 1e3 * (h1 - h0) / 3
        _MARKS = None
        phases = {n: {"gpu_ms": round(v[0], 4), "host_ms": round(v[1], 4)} for n, v in acc.items()}
    for _ in range(warm):
        step_e2e()
    ms_e2e = timed(step_e2e, steps)

    # ---- probe leg (BASELINE config 4, strong scaling): the query is cut `world` ways, windows are routed to the k-mer's
    #      owner by the same fused scatter, matched there, and the (i,j) rows are emitted on the owners
    probe = None
    if use_ranges and strong and not getattr(args, "no_probe", False):
        Lq = w["Lq"]
        q_all = synth.config_c4_query(seq_all, Lq)
        q0, q1, _, _ = shard_bounds(Lq, world, rank, k)
        q_dev = torch.from_numpy(np.ascontiguousarray(q_all[q0:q1])).to(dev)
        ixq = build(own_dev)
        rows = [None]

        def probe_step():
            rows[0] = sharded_query_ranges(ixq, q_dev, Lq, k, xchg)
        for _ in range(3):
            probe_step()
        ms_q = timed(probe_step, max(3, steps // 3))
        tot_rows = torch.tensor([rows[0].shape[0]], dtype=torch.int64, device=dev)
        dist.all_reduce(tot_rows)
        probe = {"metric": "kmers_queried_per_s", "value": (Lq - k + 1) / (ms_q * 1e-3), "unit": "k-mers/s",
                 "query_bases": Lq, "rows": int(tot_rows.item()), "ms": ms_q,
                 "what": "sharded seq.kmer.pos (c4 query cut N ways): halo, fused scatter of (mixed key, i) records over NVLink, "
                         "lookups + compaction + scan + row emission on the owners (rows stay on the owners' devices)"}
        ixq.free()
        del q_all
    sizes = torch.tensor([N, U], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    if xchg is not None:
        xchg.close()
    if rank != 0:
        return None
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    roof = None
    sp = prof.get("sort_pass_hist") or prof.get("sort_pass")
    if sp and sp[0] > 0:
        ach = sp[2] / (sp[0] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "sort_pass (rank 0)", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src, "launches": int(sp[1]),
                "avg_launch_ms": sp[0] / max(sp[1], 1), "algo_bytes_per_launch": sp[2] / max(sp[1], 1)}
    kernels = {n: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps} for n, v in sorted(prof.items())}
    exchanged = 12.0 * ntot / world * (world - 1) / world     # bytes leaving each GPU per step (uniform owners)
    sc_ms = prof.get("scatter_peer", (0, 0, 0))[0] / steps
    path = ("written straight into equal mixed-key ranges' owners over NVLink by the partitioning pass (peer memory, per-source regions: "
            "no counting, no all-to-all)" if used_ranges else
            "written straight into the key-range owners' arrays over NVLink by the partitioning pass (peer memory, no all-to-all)"
            if xchg is not None else "routed to key-range owners by one NCCL all-to-all (peer memory unavailable)")
    return {"metric": METRIC, "value": ntot / (ms * 1e-3), "unit": "k-mers/s", "n_gpus": world,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": config_of(w, k, L),
            "detail": {"sharding": (f"ONE {Ltot}-base sequence cut into {world} shards with k-1 bases of overlap" if strong else
                                    f"{world} shards of {L} bases each of one {Ltot}-base sequence") + "; (key,pos) records " + path,
                       "bases_total": Ltot, "kmers": int(ntot), "per_rank_kmers": all_sizes[:, 0].tolist(),
                       "per_rank_distinct": all_sizes[:, 1].tolist(), "l2": "inputs_exceed_l2"},
            "parity_checked": parity, "host_placement": numa, "phases_rank0": phases,
            "e2e": {"value": ntot / (ms_e2e * 1e-3), "unit": "k-mers/s", "h2d_bytes_per_step": int(Ltot),
                    "d2h_bytes_per_step": int((8 * all_sizes[:, 0] + 4 * all_sizes[:, 1]).sum()), "ms_per_step": ms_e2e,
                    "what": "per rank: pinned host shard -> device, sharded build, kmer_pos(2|8) of the owner's slice into pinned host arrays"},
            "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof, "cpu_baseline": None, "probe": probe,
            "exchange": {"bytes_per_gpu_per_step": exchanged, "what": "12-byte records leaving each GPU over NVLink (uniform owners)",
                         "kernel": "scatter_peer", "ms_per_step": sc_ms,
                         "nvlink_GBps": exchanged / (sc_ms * 1e-3) / 1e9 if sc_ms else None,
                         "nvlink_frac": exchanged / (sc_ms * 1e-3) / 1e9 / 900.0 if sc_ms else None,
                         "nvlink_peak": "900 GB/s per direction per GPU (NVLink 5)"},
            "kernels": kernels}
