"""CPU tests of the drop-in boundary: libkmergpu.so loads, exports every symbol that
include/kmergpu.h declares, and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes
import os
import re

import numpy as np

import pytest

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "kmergpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kmg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from kmer_hasher_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in kmergpu.h but not exported"
    # the ctypes table binds exactly the header's surface
    assert sorted(_lib.SIGNATURES) == syms


def test_no_cpu_fallback_and_argument_errors():
    import kmer_hasher_b200 as kh
    import torch
    # argument guards mirror the reference's messages and are raised before any device work
    with pytest.raises(ValueError, match="less than 1\\+MAX_K"):
        kh.make_kmer_hash("ACGT" * 20, 33)
    with pytest.raises(ValueError, match="at least k"):
        kh.make_kmer_hash("ACGT", 4)
    with pytest.raises(TypeError, match="external pointer"):
        kh.kmer_pos("not a pointer", 15)
    if not torch.cuda.is_available():
        with pytest.raises(kh.KmgError) as ei:
            kh.make_kmer_hash("ACGT" * 20, 4)
        assert ei.value.code in (-4, -6)


def test_product_does_not_touch_the_oracle():
    """Nothing under kmer_hasher_b200/ may import, link or execute oracle/."""
    pkg = os.path.join(ROOT, "kmer_hasher_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".c", ".h", "Makefile")):
                text = open(os.path.join(dp, fn), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), (dp, fn)
                assert "libkmer_oracle" not in text and "libkmer_ref" not in text and "kmer_oracle.c" not in text, (dp, fn)


def test_canonical_renumbering_is_a_stable_regrouping():
    """kmer_pos(..., canonical=True) re-orders a grouped index by key: rows keep their order inside a k-mer,
    k-mers are renumbered by rank.  Pure host logic (numpy), checked here on a hand-made index."""
    import kmer_hasher_b200 as kh
    # three k-mers in index order with keys 50, 10, 30 -> canonical order: 10 (was i=2), 30 (i=3), 50 (i=1)
    keys = np.array([50, 10, 30], np.uint64)
    order = np.argsort(keys, kind="stable")
    rank = np.empty(3, np.int64)
    rank[order] = np.arange(3)
    pos = np.array([[1, 7], [1, 9], [2, 3], [3, 1], [3, 4], [3, 8]], np.int32)
    got = kh._renumber(pos, rank)
    assert got.tolist() == [[1, 3], [2, 1], [2, 4], [2, 8], [3, 7], [3, 9]]
    pairs = np.array([[1, 7, 9], [3, 1, 4], [3, 1, 8], [3, 4, 8]], np.int32)
    assert kh._renumber(pairs, rank).tolist() == [[2, 1, 4], [2, 1, 8], [2, 4, 8], [3, 7, 9]]
    assert kh._renumber(np.empty((0, 2), np.int32), rank).shape == (0, 2)


def test_host_copy_of_the_staging_path():
    """kmg_host_copy (csrc/hostcopy.c): non-temporal stores above 1 MB, memcpy below; any alignment of either side."""
    import ctypes as C
    import numpy as np
    from kmer_hasher_b200 import _lib
    raw = C.CDLL(_lib.LIB_PATH)
    raw.kmg_host_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    raw.kmg_host_copy.restype = None
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, 5_000_000, dtype=np.uint8)
    for n, so, do in [(0, 0, 0), (1, 3, 5), (4096, 1, 0), ((1 << 20) - 1, 7, 9), (1 << 20, 0, 0), ((1 << 20) + 1, 1, 31), (3_000_017, 13, 1),
                      (4_194_304, 0, 33)]:
        dst = np.full(n + 128, 0xEE, np.uint8)
        raw.kmg_host_copy(dst.ctypes.data + do, src.ctypes.data + so, n)
        assert np.array_equal(dst[do:do + n], src[so:so + n])
        assert (dst[:do] == 0xEE).all() and (dst[do + n:] == 0xEE).all()      # nothing written outside
