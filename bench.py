#!/usr/bin/env python
"""bench.py -- k-mers indexed/s of the B200 k-mer position index (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3]

A step = one pass of the hot path over one synthetic sequence:
  N=1   BASELINE config 2: 40 Mbp repeat-rich sequence, k=32, make.kmer.hash + kmer.pos(2|8)
        (index + counts + positions).  `value` times it with the sequence and the outputs resident in
        HBM; `e2e` times the same calls through the public API with pinned HOST buffers (sequence in,
        pos/count matrices out), copies inside the timed region.  A probe leg (seq.kmer.pos) is
        reported beside it.
  N>1   the same per-GPU work (weak scaling): an N x 40 Mbp sequence, sharded with k-1 overlap,
        (key,pos) records routed to key-range owners by an NCCL all-to-all, per-owner sort + CSR.
--impl reference times the reference's own C (oracle/_ref, single-threaded like the reference)
on the host cores for the same metric.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

K = 32
WORKLOADS = {"c2": dict(L=40_000_000, k=32, name="c2: synthetic 40 Mbp repeat-rich, k=32, make.kmer.hash + kmer.pos(2|8)"),
             "c3": dict(L=250_000_000, k=21, name="c3: synthetic 250 Mbp with N gaps, k=21, make.kmer.hash + kmer.pos(2|8)"),
             "c3k32": dict(L=250_000_000, k=32, name="c3k32: synthetic 250 Mbp with N gaps, k=32, make.kmer.hash + kmer.pos(2|8)")}


def gen_sequence(workload: str, L: int, out=None):
    from kmer_hasher_b200 import synth
    return synth.config_c2(L, out=out) if workload == "c2" else synth.config_c3(L, out=out)


def ncu_traffic(kernel: str):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r01_sort_pass_ncu.json")
    if kernel.startswith("sort_pass") and os.path.exists(path):
        d = json.load(open(path))
        return d["traffic_bytes_per_launch"], d["source"]
    return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "app_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's own C on the host cores
# ---------------------------------------------------------------------------------------------------------
def cpu_reference(workload: str, k: int, sample_len: int, seq=None):
    """Times seq_to_hash + the kmer_positions(2|8) loop of the unmodified reference engine on a prefix."""
    from oracle import Reference
    ref = Reference()
    if seq is None:
        seq = gen_sequence(workload, sample_len)
    s = np.ascontiguousarray(seq[:sample_len])
    s = np.where(s == 0, ord("A"), s).astype(np.uint8)
    ix = ref.build(s, k)
    e = ix.extract_raw(2 | 8)
    n = ix.N
    t_build, t_ext = ix.build_seconds, e["seconds"]
    ix.close()
    return n, t_build, t_ext


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    k = w["k"]
    steps, warm = args.steps, args.warmup
    # bound each step so the whole run stays within a few minutes (~2.2 M k-mers/s single core)
    budget = 150.0 * 2.0e6
    sample = int(min(w["L"], max(2_000_000, budget / max(1, steps + warm))))
    seq = gen_sequence(args.workload, sample)
    times = []
    n = 0
    for i in range(warm + steps):
        n, tb, te = cpu_reference(args.workload, k, sample, seq)
        if i >= warm:
            times.append(tb + te)
    t = float(np.mean(times))
    val = n / t
    sample_desc = f"first {sample} bases of the {args.workload} sequence per step; seq_to_hash + kmer_positions(2|8) loop"
    line = {"impl": "reference", "metric": "kmers_indexed_per_s", "value": val, "unit": "k-mers/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": w["name"], "k": k, "sample_bases": sample},
            "cpu_baseline": {"value": val, "unit": "k-mers/s", "cores": 1, "kind": "reference", "sample": sample_desc},
            "e2e": {"value": val, "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import kmer_hasher_b200 as kh
    from kmer_hasher_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")
    torch.cuda.set_device(local)
    _lib.check(_lib.load().kmg_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    _lib.check(_lib.load().kmg_set_stream(stream.cuda_stream))

    w = WORKLOADS[args.workload]
    k, L = w["k"], w["L"]
    steps, warm = args.steps, max(args.warmup, 3)
    hbm_peak, peak_src = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world == 1:
        line = bench_single(args, kh, torch, w, k, L, steps, warm, hbm_peak, peak_src, local)
    else:
        from kmer_hasher_b200 import dist as kdist
        line = kdist.bench_sharded(args, w, k, L, steps, warm, hbm_peak, peak_src, barrier)
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_single(args, kh, torch, w, k, L, steps, warm, hbm_peak, peak_src, dev):
    # ---- inputs: pinned host copy and a device-resident copy -------------------------------------------
    seq_pin = kh.pinned_empty(L, np.uint8)
    gen_sequence(args.workload, L, out=seq_pin)
    seq_dev = torch.from_numpy(np.asarray(seq_pin)).cuda()
    ix = kh.make_kmer_hash(seq_dev, k)
    U, N, P = ix.sizes
    ix.free()
    pos_dev = torch.empty((N, 2), dtype=torch.int32, device="cuda")
    cnt_dev = torch.empty(U, dtype=torch.int32, device="cuda")
    pos_pin = kh.pinned_empty((N, 2), np.int32)
    cnt_pin = kh.pinned_empty(U, np.int32)

    def step_device():
        h = kh.make_kmer_hash(seq_dev, k)
        kh.kmer_pos(h, 2 | 8, out={"pos": pos_dev, "count": cnt_dev})
        h.free()

    def step_e2e():
        t0 = time.perf_counter()
        h = kh.make_kmer_hash(seq_pin, k)
        t1 = time.perf_counter()
        kh.kmer_pos(h, 2 | 8, out={"pos": pos_pin, "count": cnt_pin})
        t2 = time.perf_counter()
        h.free()
        if os.environ.get("KMG_BENCH_DEBUG"):
            print("e2e step: build %.2f extract %.2f free %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3), file=sys.stderr)

    def timed(fn, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    for _ in range(warm):
        step_device()
    kh.profile(enable=True, reset=True)
    kh.profile(reset=True)
    l0 = kh.launch_count()
    with ClockSampler(dev) as clk:
        ms = timed(step_device, steps)
    launches = kh.launch_count() - l0
    prof = kh.profile(enable=False)
    kh.profile(reset=True)

    for _ in range(warm):
        step_e2e()
    ms_e2e = timed(step_e2e, steps)

    # ---- probe leg (seq.kmer.pos): 10 Mbp query against the same index ------------------------------------
    probe = None
    if not args.no_probe:
        from kmer_hasher_b200 import synth
        Lq = min(10_000_000, L // 4)
        q = synth.generate(Lq, 0xC4)                               # unrelated random background
        srcseq = np.asarray(seq_pin)
        rng = np.random.default_rng(4)
        for _ in range(Lq // 50_000):                              # sprinkle 2 kb copies of index sequence
            a, b = int(rng.integers(0, L - 2000)), int(rng.integers(0, Lq - 2000))
            q[b:b + 2000] = srcseq[a:a + 2000]
        q_dev = torch.from_numpy(q).cuda()
        h = kh.make_kmer_hash(seq_dev, k)
        import ctypes as C
        Lb = _libmod().load()
        st, M = C.c_void_p(), C.c_uint64()

        def probe_count():
            _libmod().check(Lb.kmg_query_begin(h._handle(), q_dev.data_ptr(), Lq, k, C.byref(st), C.byref(M)))
            Lb.kmg_query_free(st)

        for _ in range(3):
            probe_count()
        ms_q = timed(probe_count, max(3, steps // 3))
        probe = {"metric": "kmers_queried_per_s", "value": (Lq - k + 1) / (ms_q * 1e-3), "unit": "k-mers/s",
                 "query_bases": Lq, "rows": int(M.value), "ms": ms_q, "what": "kmg_query_begin (encode+match+scan), device-resident"}
        h.free()

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------
    dom = max(prof.items(), key=lambda kv: kv[1][0]) if prof else None
    roof = None
    if dom:
        name, (tms, nl, bytes_) = dom
        ach = bytes_ / (tms * 1e-3) / 1e9 if tms > 0 else 0.0
        traffic, traffic_src = ncu_traffic(name)
        roof = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launches": int(nl),
                "avg_launch_ms": tms / max(nl, 1), "algo_bytes_per_launch": bytes_ / max(nl, 1)}
    kernels = {n: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps,
                   "GBps": (v[2] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else None)} for n, v in sorted(prof.items())}

    # ---- CPU baseline beside it -----------------------------------------------------------------------------
    cpu = None
    if not args.no_cpu:
        sample = min(L, args.cpu_sample)
        n_cpu, tb, te = cpu_reference(args.workload, k, sample, np.asarray(seq_pin))
        cpu = {"value": n_cpu / (tb + te), "unit": "k-mers/s", "cores": 1, "kind": "reference",
               "sample": f"first {sample} bases of the same sequence; reference seq_to_hash {tb:.2f}s + kmer_positions(2|8) loop {te:.2f}s, "
                         f"gcc -O2, 1 thread (the reference path is single-threaded); host has {os.cpu_count()} cores",
               "build_only_value": n_cpu / tb}

    # whole step against SURVEY.md 8d's accounting (an 8-bit LSD sort by key: R = ceil(2k/8) passes) and against
    # the bytes the shipped step really needs (grouped build: 5 passes + one detection read of the records)
    R_key = (2 * k + 7) // 8
    R_used = 5 if R_key > 6 else R_key
    survey_bytes = L + (36 + 24 * R_key) * N + 20 * U + (12 * U + 12 * N) + 8 * U
    used_bytes = L + L + (12 + 24 * (R_used - 1)) * N + (8 * N if R_used != R_key else 0) + (8 * N + 12 * U) + 4 * U + (12 * U + 12 * N) + 8 * U
    step_roof = {"bound": "hbm", "what": "build + kmer.pos(2|8), all kernels of the step",
                 "survey_formula_bytes": int(survey_bytes), "survey_formula_GBps": survey_bytes / (ms * 1e-3) / 1e9,
                 "survey_formula_frac": survey_bytes / (ms * 1e-3) / 1e9 / hbm_peak,
                 "algorithmic_bytes_as_built": int(used_bytes), "achieved": used_bytes / (ms * 1e-3) / 1e9,
                 "frac": used_bytes / (ms * 1e-3) / 1e9 / hbm_peak, "peak": hbm_peak, "unit": "GB/s",
                 "note": "survey_formula_* charge the 8 key passes of SURVEY.md 8d although the grouped build runs 5"}
    h2d = L
    d2h = 8 * N + 4 * U
    return {"metric": "kmers_indexed_per_s", "value": N / (ms * 1e-3), "unit": "k-mers/s", "n_gpus": 1, "steps": steps,
            "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": w["name"], "k": k, "bases": L, "kmers": int(N), "distinct": int(U),
                       "kmer_order": "grouped (make.kmer.hash default for k >= 25: 5 radix passes on a mix of the key; "
                                     "do.sort=TRUE would give ascending keys in 8)" if k >= 25 else "ascending key",
                       "l2": "inputs_exceed_l2 (keys 8N + pos 4N bytes per pass >> 126 MB)"},
            "e2e": {"value": N / (ms_e2e * 1e-3), "unit": "k-mers/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
                    "what": "make_kmer_hash(pinned host seq) + kmer_pos(2|8) into pinned host arrays"},
            "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roof, "step_roofline": step_roof,
            "cpu_baseline": cpu, "probe": probe, "kernels": kernels}


def _libmod():
    from kmer_hasher_b200 import _lib
    return _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=40_000_000)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
