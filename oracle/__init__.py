"""oracle -- TEST INFRASTRUCTURE ONLY (ctypes front ends of the two CPU checkers).

* ``Oracle``    : oracle/_build/libkmer_oracle.so, the plain-C restatement
                  (oracle/kmer_oracle.c).  Travels to the GPU box.
* ``Reference`` : oracle/_ref/libkmer_ref.so, the UNMODIFIED reference engine
                  (/root/reference/src/kmer_pos.c + kmer_util.c) behind
                  oracle/ref_driver.c.  Built only where /root/reference exists;
                  the prebuilt .so travels to the GPU box.

Only tests/, bench.py's cpu_baseline / ``--impl reference`` leg and
``__graft_entry__.smoke()`` may import this package.  Nothing under
``kmer_hasher_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_build", "libkmer_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libkmer_ref.so")

_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_dblp = C.POINTER(C.c_double)


def build(quiet: bool = True) -> None:
    """Compile the checkers (``make -C oracle``). Building is not using."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


def _as_bytes(seq) -> bytes:
    if isinstance(seq, str):
        return seq.encode("latin-1")
    if isinstance(seq, np.ndarray):
        return seq.tobytes()
    return bytes(seq)


class OracleIndex:
    """Canonical CSR index made by the plain-C restatement."""

    def __init__(self, lib, handle, k):
        self._lib, self._h, self.k = lib, handle, k
        U, N, P = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib.ko_sizes(handle, C.byref(U), C.byref(N), C.byref(P))
        self.U, self.N, self.P = U.value, N.value, P.value

    def extract(self, flag: int = 15, with_keys: bool = True):
        """dict(keys, kmer, pos, pair_pos, count) in canonical order."""
        k = self.k
        keys = np.empty(self.U, np.uint64) if with_keys else None
        kmers = np.empty(self.U * (k + 1), np.uint8) if flag & 1 else None
        pos = np.empty(2 * self.N, np.int32) if flag & 2 else None
        pairs = np.empty(3 * self.P, np.int32) if flag & 4 else None
        counts = np.empty(self.U, np.int32) if flag & 8 else None
        self._lib.ko_extract(self._h, _ptr(keys, _u64p), _ptr(kmers, C.c_char_p), _ptr(pos, _i32p),
                             _ptr(pairs, _i32p), _ptr(counts, _i32p))
        return dict(keys=keys, kmer=kmers, pos=pos, pair_pos=pairs, count=counts)

    def query(self, seq, k: int) -> np.ndarray:
        b = _as_bytes(seq)
        rows = _i32p()
        n = self._lib.ko_query(self._h, b, len(b), k, C.byref(rows))
        out = np.ctypeslib.as_array(rows, shape=(2 * n,)).copy() if n else np.empty(0, np.int32)
        if n:
            self._lib.ko_free_buf(rows)
        return out

    def query_count(self, seq, k: int) -> int:
        b = _as_bytes(seq)
        return self._lib.ko_query(self._h, b, len(b), k, None)

    def close(self):
        if self._h:
            self._lib.ko_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def pairs_join(a: dict, b: dict) -> np.ndarray:
    """kmer.pairs restated (kmer_pair_pos, src/kmer_hash.c:1174-1203) from two canonical extractions
    (`extract(2 | 8)` with keys): for every k-mer of `a` that `b` also holds, rows (a_pos, b_pos), a
    position outer, b position inner (:1190-1195), a's k-mers in ascending key order.  Flattened."""
    pa, pb = a["pos"].reshape(-1, 2)[:, 1], b["pos"].reshape(-1, 2)[:, 1]
    sa = np.concatenate([[0], np.cumsum(a["count"], dtype=np.int64)])
    sb = np.concatenate([[0], np.cumsum(b["count"], dtype=np.int64)])
    idx = np.searchsorted(b["keys"], a["keys"])
    idx[idx >= len(b["keys"])] = 0
    shared = np.nonzero(b["keys"][idx] == a["keys"])[0] if len(b["keys"]) else np.empty(0, np.int64)
    out = []
    for u in shared:
        v = idx[u]
        la, lb = pa[sa[u]:sa[u + 1]], pb[sb[v]:sb[v + 1]]
        out.append(np.stack([np.repeat(la, len(lb)), np.tile(lb, len(la))], axis=1))
    return (np.concatenate(out) if out else np.empty((0, 2), np.int32)).astype(np.int32).ravel()



class _RefDig(C.Structure):
    _fields_ = [("n", C.c_uint64), ("sum", C.c_uint64), ("wsum", C.c_uint64)]

    def tuple(self):
        return (int(self.n), int(self.sum), int(self.wsum))


def digest(values) -> tuple:
    """(n, sum v_t, sum v_t * (2t + 1)) mod 2^64 of a flattened integer stream: the digest ref_driver.c's
    ref_digest_canonical / ref_query_digest record for the full-size golden fixtures (numpy restatement)."""
    v = np.ascontiguousarray(values).ravel()
    v = v.astype(np.int64).view(np.uint64) if v.dtype != np.uint64 else v
    t = np.arange(v.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return (int(v.size), int(v.sum(dtype=np.uint64)), int((v * (np.uint64(2) * t + np.uint64(1))).sum(dtype=np.uint64)))


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        lib = C.CDLL(ORACLE_SO)
        lib.ko_windows.restype = C.c_int64
        lib.ko_windows.argtypes = [C.c_char_p, C.c_int64, C.c_int, _u64p, _i32p]
        lib.ko_build.restype = C.c_void_p
        lib.ko_build.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.ko_build_from_records.restype = C.c_void_p
        lib.ko_build_from_records.argtypes = [_u64p, _i32p, C.c_int64, C.c_int]
        lib.ko_free.argtypes = [C.c_void_p]
        lib.ko_sizes.argtypes = [C.c_void_p, _u64p, _u64p, _u64p]
        lib.ko_extract.argtypes = [C.c_void_p, _u64p, C.c_char_p, _i32p, _i32p, _i32p]
        lib.ko_query.restype = C.c_int64
        lib.ko_query.argtypes = [C.c_void_p, C.c_char_p, C.c_int64, C.c_int, C.POINTER(_i32p)]
        lib.ko_free_buf.argtypes = [C.c_void_p]
        self.lib = lib

    def windows(self, seq, k: int):
        """(keys, pos) stream in emission order."""
        b = _as_bytes(seq)
        cap = max(len(b) - k + 1, 1)
        keys = np.empty(cap, np.uint64)
        pos = np.empty(cap, np.int32)
        n = self.lib.ko_windows(b, len(b), k, _ptr(keys, _u64p), _ptr(pos, _i32p))
        return keys[:n].copy(), pos[:n].copy()

    def build(self, seq, k: int, guard: bool = True) -> OracleIndex:
        b = _as_bytes(seq)
        err = C.c_int(0)
        h = self.lib.ko_build(b, len(b), k, int(guard), C.byref(err))
        if not h:
            raise ValueError({1: "k must be a positive integer less than 1+MAX_K",
                              2: "the length of the sequence must be at least k"}[err.value])
        return OracleIndex(self.lib, h, k)

    def build_from_records(self, keys: np.ndarray, pos: np.ndarray, k: int) -> OracleIndex:
        keys = np.ascontiguousarray(keys, np.uint64)
        pos = np.ascontiguousarray(pos, np.int32)
        h = self.lib.ko_build_from_records(_ptr(keys, _u64p), _ptr(pos, _i32p), len(keys), k)
        return OracleIndex(self.lib, h, k)


class ReferenceIndex:
    """An index held by the reference's own khash (through ref_driver.c)."""

    def __init__(self, lib, handle, k, seconds):
        self._lib, self._h, self.k, self.build_seconds = lib, handle, k, seconds
        U, N, P, B, W = (C.c_uint64() for _ in range(5))
        lib.ref_sizes(handle, C.byref(U), C.byref(N), C.byref(P), C.byref(B), C.byref(W))
        self.U, self.N, self.P, self.buckets, self.new_kmers = U.value, N.value, P.value, B.value, W.value

    def _alloc(self, flag, with_keys=True):
        k = self.k
        keys = np.empty(self.U, np.uint64) if with_keys else None
        kmers = np.empty(self.U * (k + 1), np.uint8) if flag & 1 else None
        pos = np.empty(2 * self.N, np.int32) if flag & 2 else None
        pairs = np.empty(3 * self.P, np.int32) if flag & 4 else None
        counts = np.empty(self.U, np.int32) if flag & 8 else None
        return keys, kmers, pos, pairs, counts

    def extract_raw(self, flag: int = 15):
        """kmer.pos exactly as R would receive it (khash bucket order) + seconds."""
        keys, kmers, pos, pairs, counts = self._alloc(flag)
        sec = C.c_double(0)
        self._lib.ref_extract_raw(self._h, _ptr(keys, _u64p), _ptr(kmers, C.c_char_p), _ptr(pos, _i32p),
                                  _ptr(pairs, _i32p), _ptr(counts, _i32p), C.byref(sec))
        return dict(keys=keys, kmer=kmers, pos=pos, pair_pos=pairs, count=counts, seconds=sec.value)

    def extract(self, flag: int = 15):
        """kmer.pos in canonical order (k-mers by ascending key, i remapped)."""
        keys, kmers, pos, pairs, counts = self._alloc(flag)
        self._lib.ref_extract_canonical(self._h, _ptr(keys, _u64p), _ptr(kmers, C.c_char_p),
                                        _ptr(pos, _i32p), _ptr(pairs, _i32p), _ptr(counts, _i32p))
        return dict(keys=keys, kmer=kmers, pos=pos, pair_pos=pairs, count=counts)

    def query(self, seq, k: int, want_rows: bool = True):
        b = _as_bytes(seq)
        rows = _i32p()
        sec = C.c_double(0)
        n = self._lib.ref_query(self._h, b, k, C.byref(rows), C.byref(sec))
        self.query_seconds = sec.value
        out = None
        if want_rows:
            out = np.ctypeslib.as_array(rows, shape=(2 * n,)).copy() if n else np.empty(0, np.int32)
        if rows:
            self._lib.ref_query_free(rows)
        return out if want_rows else n

    def digest(self, flag: int = 2 | 8) -> dict:
        """Digests (see `digest`) of the canonical keys / count / pos / pair.pos streams, generated from the
        reference's tables without materialising them (ref_digest_canonical)."""
        d = {n: _RefDig() for n in ("keys", "count", "pos", "pair_pos")}
        bind = (C.c_uint64 * 2)()
        self._lib.ref_digest_canonical(self._h, C.byref(d["keys"]), C.byref(d["count"]) if flag & 8 else None,
                                       C.byref(d["pos"]) if flag & 2 else None, C.byref(d["pair_pos"]) if flag & 4 else None, bind)
        out = {n: v.tuple() for n, v in d.items() if n == "keys" or (n == "count" and flag & 8) or (n == "pos" and flag & 2)
               or (n == "pair_pos" and flag & 4)}
        out["bind"] = (int(bind[0]), int(bind[1]))    # order-independent: sum key*count, sum key*pos (mod 2^64)
        return out

    def query_digest(self, seq, k: int):
        """(rows, digest of the interleaved (i,j) stream) of seq_kmer_positions."""
        b = _as_bytes(seq)
        d, sec = _RefDig(), C.c_double(0)
        n = self._lib.ref_query_digest(self._h, b, k, C.byref(d), C.byref(sec))
        self.query_seconds = sec.value
        return int(n), d.tuple()

    def pairs_join(self, other: "ReferenceIndex") -> np.ndarray:
        """kmer.pairs(self, other) on the reference's own hash tables (ref_pairs_join): rows (a, b) flattened."""
        rows = _i32p()
        n = self._lib.ref_pairs_join(self._h, other._h, C.byref(rows))
        out = np.ctypeslib.as_array(rows, shape=(2 * n,)).copy() if n else np.empty(0, np.int32)
        self._lib.ref_free_buf(rows)
        return out

    def close(self):
        if self._h:
            self._lib.ref_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


class Reference:
    def __init__(self):
        if not os.path.exists(REF_SO):
            if os.path.isdir("/root/reference/src"):
                build()
            else:
                raise FileNotFoundError(f"{REF_SO} missing and /root/reference absent")
        lib = C.CDLL(REF_SO)
        lib.ref_build.restype = C.c_void_p
        lib.ref_build.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int), _dblp]
        lib.ref_build_core.restype = C.c_void_p
        lib.ref_build_core.argtypes = [C.c_char_p, C.c_int]
        lib.ref_free.argtypes = [C.c_void_p]
        lib.ref_sizes.argtypes = [C.c_void_p, _u64p, _u64p, _u64p, _u64p, _u64p]
        lib.ref_extract_raw.argtypes = [C.c_void_p, _u64p, C.c_char_p, _i32p, _i32p, _i32p, _dblp]
        lib.ref_extract_canonical.argtypes = [C.c_void_p, _u64p, C.c_char_p, _i32p, _i32p, _i32p]
        lib.ref_query.restype = C.c_int64
        lib.ref_query.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(_i32p), _dblp]
        lib.ref_query_free.argtypes = [_i32p]
        lib.ref_window_stream.restype = C.c_int64
        lib.ref_window_stream.argtypes = [C.c_char_p, C.c_int, C.POINTER(_u64p), C.POINTER(_i32p)]
        lib.ref_free_buf.argtypes = [C.c_void_p]
        lib.ref_pairs_join.restype = C.c_int64
        lib.ref_pairs_join.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_i32p)]
        lib.ref_reads_parse.restype = C.c_int64
        lib.ref_reads_parse.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_int64))]
        lib.ref_reads_free.argtypes = [C.c_void_p]
        lib.ref_count_new.restype = C.c_void_p
        lib.ref_count_new.argtypes = [C.c_int]
        lib.ref_count_add.restype = C.c_int
        lib.ref_count_add.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
        lib.ref_digest_canonical.restype = None
        lib.ref_digest_canonical.argtypes = [C.c_void_p] + [C.POINTER(_RefDig)] * 4 + [_u64p]
        lib.ref_query_digest.restype = C.c_int64
        lib.ref_query_digest.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(_RefDig), _dblp]
        self.lib = lib

    def count_kmers(self, seqs, k: int, source: int, source_n: int, table: "ReferenceIndex | None" = None) -> "ReferenceIndex":
        """count.kmers(seq, c(k, source, source_n), hash.ptr) (kmer_hash.R:43-46 -> count_kmers, src/kmer_hash.c:548-591):
        restated seq_to_counts on the reference's own khash (ref_count_add).  Read the table with extract():
        'pos' rows are (i, count of source 0), (i, count of source 1), ... and 'count' is source_n everywhere."""
        if k < 1 or k > 32:
            raise ValueError("k must be a positive integer less than 1+MAX_K")
        if source_n < 1 or source >= source_n:
            raise ValueError("source_n must be larger than 1 and larger than source")
        if table is None:
            table = ReferenceIndex(self.lib, self.lib.ref_count_new(k), k, 0.0)
        elif table.k != k:
            raise ValueError("mismatch between specified k and that given in the external pointer")
        if isinstance(seqs, (str, bytes, np.ndarray)):
            seqs = [seqs]
        for sq in seqs:
            self.lib.ref_count_add(table._h, _as_bytes(sq), source, source_n)
        U, N, P, B, W = (C.c_uint64() for _ in range(5))
        self.lib.ref_sizes(table._h, C.byref(U), C.byref(N), C.byref(P), C.byref(B), C.byref(W))
        table.U, table.N, table.P, table.buckets, table.new_kmers = U.value, N.value, P.value, B.value, W.value
        return table

    def parse_reads(self, path: str):
        """(names, sequences) of a FASTA/FASTQ file (plain or gz) as the reference's own parser reads it: klib kseq.h over
        zlib, unmodified (oracle/ref_fasta.c; src/kmer_reader.c:41-77 is the reference's read loop)."""
        names, seqs, lens = C.c_char_p(), C.c_void_p(), C.POINTER(C.c_int64)()
        n = self.lib.ref_reads_parse(path.encode(), C.byref(names), C.byref(seqs), C.byref(lens))
        if n < 0:
            raise FileNotFoundError(path)
        out_names = names.value.decode().split("\n")[:n] if n else []
        ll = [lens[i] for i in range(n)]
        raw = C.string_at(seqs, sum(ll) + n) if n else b""
        out_seqs, o = [], 0
        for ln in ll:
            out_seqs.append(raw[o:o + ln]); o += ln + 1
        for p_ in (names, seqs, lens):
            self.lib.ref_reads_free(p_)
        return out_names, out_seqs

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO) or os.path.isdir("/root/reference/src")

    def build(self, seq, k: int, do_sort: bool = False, guard: bool = True) -> ReferenceIndex:
        b = _as_bytes(seq)  # ctypes appends the NUL the reference scans for
        assert b"\0" not in b
        if guard:
            err, sec = C.c_int(0), C.c_double(0)
            h = self.lib.ref_build(b, k, int(do_sort), C.byref(err), C.byref(sec))
            if not h:
                raise ValueError({1: "k must be a positive integer less than 1+MAX_K",
                                  2: "the length of the sequence must be at least k"}[err.value])
            return ReferenceIndex(self.lib, h, k, sec.value)
        return ReferenceIndex(self.lib, self.lib.ref_build_core(b, k), k, 0.0)

    def windows(self, seq, k: int):
        b = _as_bytes(seq)
        kp, pp = _u64p(), _i32p()
        n = self.lib.ref_window_stream(b, k, C.byref(kp), C.byref(pp))
        keys = np.ctypeslib.as_array(kp, shape=(max(n, 1),))[:n].copy()
        pos = np.ctypeslib.as_array(pp, shape=(max(n, 1),))[:n].copy()
        self.lib.ref_free_buf(kp)
        self.lib.ref_free_buf(pp)
        return keys, pos
