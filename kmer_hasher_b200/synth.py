"""Synthetic inputs for the BASELINE.json configurations (ctypes front end of csrc/synth.c)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "csrc", "libkmersynth.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "libkmersynth.so"], check=True,
                           stdout=subprocess.DEVNULL)
        lib = C.CDLL(_SO)
        lib.kms_generate.restype = C.c_int
        lib.kms_generate.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.POINTER(C.c_double), C.c_int]
        lib.kms_make_query.restype = C.c_int
        lib.kms_make_query.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_uint64, C.c_double,
                                       C.c_double, C.c_double, C.c_int64, C.c_int64]
        _lib = lib
    return _lib


def generate(L: int, seed: int, *, repeat=0.0, tandem=0.0, homo=0.0, lower=0.0, n_gaps=0, gap_min=10,
             gap_max=500000, n_single=0, tail_k=0, tandem_unit_max=60, tandem_len_max=200000,
             n_families=200, tandem_len_min=1000, homo_len_max=500, out: np.ndarray | None = None) -> np.ndarray:
    """A deterministic synthetic sequence as a uint8 array of ASCII bases."""
    if out is None:
        out = np.empty(L, np.uint8)
    assert out.dtype == np.uint8 and out.size >= L and out.flags.c_contiguous
    p = (C.c_double * 14)(repeat, tandem, homo, lower, n_gaps, gap_min, gap_max, n_single, tail_k,
                          tandem_unit_max, tandem_len_max, n_families, tandem_len_min, homo_len_max)
    rc = _load().kms_generate(out.ctypes.data, L, seed, p, 14)
    if rc:
        raise ValueError("kms_generate failed")
    return out[:L]


def make_query(ref: np.ndarray, Lq: int, seed: int, *, unrelated=0.2, sub=0.01, indel=0.001, seg_min=10_000,
               seg_max=5_000_000, out: np.ndarray | None = None) -> np.ndarray:
    if out is None:
        out = np.empty(Lq, np.uint8)
    ref = np.ascontiguousarray(ref, np.uint8)
    rc = _load().kms_make_query(ref.ctypes.data, ref.size, out.ctypes.data, Lq, seed, unrelated, sub, indel,
                                seg_min, seg_max)
    if rc:
        raise ValueError("kms_make_query failed")
    return out[:Lq]


# ---- the BASELINE.json configurations, scalable by `scale` for tests -----------------
def config_c2(L: int = 40_000_000, out=None, seed: int = 0xC2) -> np.ndarray:
    """40 Mbp repeat-rich, no N (index at k=32)."""
    return generate(L, seed, repeat=0.30, tandem=0.10, homo=0.05, lower=0.20, out=out)


def config_c3(L: int = 250_000_000, tail_k: int = 0, out=None, seed: int = 0xC3) -> np.ndarray:
    """250 Mbp chromosome with N gaps (index at k=21; also C4's index at k=32).  Tandem arrays and
    homopolymer runs are kept short (microsatellite scale) so that the C4 dot plot stays below the
    2^31-1 rows an R matrix can hold; config 2 carries the heavy satellite arrays instead."""
    scale = L / 250_000_000
    return generate(L, seed, repeat=0.30, tandem=0.02, homo=0.0005, lower=0.20, tandem_len_min=100,
                    tandem_len_max=2000, homo_len_max=60,
                    n_gaps=max(1, int(60 * scale)), gap_min=10, gap_max=max(10, int(500_000 * scale)),
                    n_single=max(1, int(2000 * scale)), tail_k=tail_k, out=out)


def config_c4_query(index_seq: np.ndarray, Lq: int = 100_000_000, out=None) -> np.ndarray:
    """100 Mbp query made of diverged copies of index segments + 20 % unrelated sequence."""
    scale = Lq / 100_000_000
    return make_query(index_seq, Lq, 0xC4, unrelated=0.2, sub=0.01, indel=0.001,
                      seg_min=max(100, int(10_000 * scale)), seg_max=max(1000, int(5_000_000 * scale)), out=out)


def config_c5(L: int = 40_000_000, out=None) -> np.ndarray:
    """40 Mbp tandem-repeat-heavy: 1e9 < pairs < 2^31 at k=12 (tuned with the oracle)."""
    return generate(L, 0xC5, tandem=C5_TANDEM_FRAC * (40_000_000 / L if L < 40_000_000 else 1.0),
                    tandem_unit_max=40, tandem_len_max=60_000, out=out)


C5_TANDEM_FRAC = 0.035
