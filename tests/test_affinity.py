"""Host placement helper (kmer_hasher_b200/affinity.py): parsing and the no-information paths (CPU only)."""
import os

from kmer_hasher_b200 import affinity


def test_parse_cpulist():
    assert affinity._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert affinity._parse_cpulist("") == set()
    assert affinity._parse_cpulist("5") == {5}


def test_unknown_device_changes_nothing():
    before = os.sched_getaffinity(0)
    cpus, node = affinity.gpu_cpus("ffff:ff:1f.7")         # no such PCI device
    assert cpus == set() and node is None
    info = affinity.bind_near_gpu(0)                        # no CUDA device here: reports why, binds nothing
    assert info["bound"] is False and "why" in info
    assert os.sched_getaffinity(0) == before
