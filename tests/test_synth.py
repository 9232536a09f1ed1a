"""CPU tests of the synthetic-input generator: the oracle, the reference engine and the GPU must all see
identical bytes on every machine, so the generator's output is pinned by checksum."""
import numpy as np

from conftest import sha
from kmer_hasher_b200 import synth


def test_generator_is_deterministic():
    assert sha(synth.config_c2(200_000)) == "8cc253205f9a8563593a743dcba8186171fadaf3ef8723e0703507348eb521f2"
    assert sha(synth.config_c3(300_000, tail_k=21)) == "8ffbb23a52c29284f3effdf8324557d4240ad78bd1ad669c86194a86fefe9e06"
    assert sha(synth.config_c5(100_000)) == "26685a151105a887fe85124dcd30641b117b43671d789111823f16d9deb531ad"
    s = synth.config_c3(300_000)
    assert sha(synth.config_c4_query(s, 50_000)) == "55e57b98410e510d031db0bb7f919902850c5ff3bb01cfbe3a1a0965f1b46756"


def test_generator_content():
    s = synth.config_c3(400_000, tail_k=21)
    assert s.dtype == np.uint8 and len(s) == 400_000
    assert set(np.unique(s | 0x20)) <= set(b"acgtn")            # only bases and N, both cases
    assert ((s | 0x20) == ord("n")).sum() > 0 and (s >= ord("a")).sum() > 0
    assert (s[-22] | 0x20) == ord("n") and not ((s[-21:] | 0x20) == ord("n")).any()   # final run of exactly k
    c2 = synth.config_c2(300_000)
    assert not ((c2 | 0x20) == ord("n")).any()                  # config 2 has no N
    out = np.empty(1000, np.uint8)
    assert synth.generate(1000, 5, out=out).base is out or np.shares_memory(out, synth.generate(1000, 5, out=out))


def test_c5_is_tandem_heavy(oracle):
    s = synth.generate(300_000, 0xC5, tandem=0.035, tandem_unit_max=40, tandem_len_max=6000)
    ix = oracle.build(s, 12)
    assert ix.P > 5 * ix.N                                       # pairs dominate: the point of config 5
