"""2+ GPU check of the fused partition + exchange (run under torchrun): the peer-memory build must
equal the general (NCCL all-to-all) build on every rank; both are timed.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/p2p_check.py [L_per_rank] [k]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import kmer_hasher_b200 as kh
from kmer_hasher_b200 import dist as kdist, synth

import faulthandler
faulthandler.dump_traceback_later(int(os.environ.get("P2P_WATCHDOG", "70")), exit=True)   # a hang prints where every rank is
Lr = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 32
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world, rank = dist.get_world_size(), dist.get_rank()
eng = kdist.CudaEngine(dev)
own = torch.from_numpy(synth.generate(Lr, 0xC2 + 7919 * rank, repeat=0.30, tandem=0.10, homo=0.05, lower=0.20, n_gaps=3, gap_max=5000, n_single=20)).to(dev)
L = Lr * world
xchg = kdist.PeerExchange(eng, int(Lr * 1.25) + 1024)


def timed(fn, n=10):
    for _ in range(3):
        fn().free()
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn().free()
    b.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# parity on a quarter-size sequence: both paths' per-owner slices, concatenated over ranks, must be identical
# (the two paths sample differently, so their splitters and per-rank slices differ)
Lq4 = Lr // 4
own4 = own[:Lq4].contiguous()
Lqq = 200_000
q = torch.from_numpy(synth.generate(Lqq, 0xC4 + rank)).to(dev)          # random background + one 5 kb copy of the index
q[1000:6000] = torch.from_numpy(synth.generate(5000, 0xABC)).to(dev)
own4[70000:75000] = q[1000:6000]
for name, ix in (("general", kdist.sharded_build(own4, Lq4 * world, k, eng)), ("peer", kdist.sharded_build_p2p(own4, Lq4 * world, k, eng, xchg))):
    e = kh.kmer_pos(ix.local, 2 | 8)
    np.save(f"/tmp/p2p_{name}_{rank}_keys.npy", kh.kmer_keys(ix.local))
    np.save(f"/tmp/p2p_{name}_{rank}_count.npy", e["count"])
    np.save(f"/tmp/p2p_{name}_{rank}_pos.npy", e["pos"][:, 1].copy())
    if name == "peer":
        b = ix
        rb = kdist.sharded_query_p2p(ix, q, Lqq * world, k, xchg).cpu().numpy()
    else:
        ra = kdist.sharded_query(ix, q, Lqq * world, k).cpu().numpy()
    np.save(f"/tmp/p2p_{name}_{rank}_rows.npy", ra if name == "general" else rb)
    N_all = ix.N_all
    ix.local.free()
dist.barrier()
if rank == 0:
    def canon(name):
        keys = np.concatenate([np.load(f"/tmp/p2p_{name}_{r}_keys.npy") for r in range(world)])
        cnt = np.concatenate([np.load(f"/tmp/p2p_{name}_{r}_count.npy") for r in range(world)])
        pos = np.concatenate([np.load(f"/tmp/p2p_{name}_{r}_pos.npy") for r in range(world)])
        start = np.concatenate([[0], np.cumsum(cnt, dtype=np.int64)])
        o = np.argsort(keys, kind="stable")             # the peer path builds grouped indexes: order by key to compare
        lists = np.concatenate([pos[start[u]:start[u + 1]] for u in o]) if len(o) < 3_000_000 else None
        return keys[o], cnt[o], lists, pos
    gk, gc, gl, gp = canon("general")
    pk, pc, pl, pp = canon("peer")
    assert np.array_equal(gk, pk) and np.array_equal(gc, pc)
    if gl is not None:
        assert np.array_equal(gl, pl)
    else:                                               # large: the multiset of positions and every list ascending
        assert np.array_equal(np.sort(gp), np.sort(pp))
    g = np.concatenate([np.load(f"/tmp/p2p_general_{r}_rows.npy") for r in range(world)])
    p = np.concatenate([np.load(f"/tmp/p2p_peer_{r}_rows.npy") for r in range(world)])
    g = g[np.lexsort((g[:, 1], g[:, 0]))]; p = p[np.lexsort((p[:, 1], p[:, 0]))]
    assert np.array_equal(g, p), "query rows"
    rb = p
b = type("B", (), {"N_all": N_all})()
if rank == 0:
    print(f"parity ok: world {world}, N_all {b.N_all}, query rows on rank 0: {len(rb)}", flush=True)
t_a = timed(lambda: kdist.sharded_build(own, L, k, eng).local)
t_b = timed(lambda: kdist.sharded_build_p2p(own, L, k, eng, xchg).local)
if rank == 0:
    n = world * (Lr - k + 1)
    print(f"general (all-to-all) build {t_a:.3f} ms = {n / t_a / 1e6:.2f} G k-mers/s | peer-memory build {t_b:.3f} ms = {n / t_b / 1e6:.2f} G k-mers/s", flush=True)
kh.profile(enable=True, reset=True)
kh.profile(reset=True)
for _ in range(5):
    kdist.sharded_build_p2p(own, L, k, eng, xchg).local.free()
torch.cuda.synchronize()
prof = kh.profile(enable=False)
if rank == 0:
    tot = 0.0
    for name, v in sorted(prof.items()):
        print(f"    {name:18s} {v[0] / 5:7.3f} ms/build  {v[1] / 5:4.1f} launches", flush=True)
        tot += v[0] / 5
    print(f"    kernel time per build {tot:.3f} ms", flush=True)
kdist._MARKS = []
for _ in range(3):
    del kdist._MARKS[:]
    kdist.sharded_build_p2p(own, L, k, eng, xchg).local.free()
torch.cuda.synchronize()
if rank == 0:
    m = kdist._MARKS
    for (n0, e0, h0), (n1, e1, h1) in zip(m[:-1], m[1:]):
        print(f"    phase {n1:10s} gpu {e0.elapsed_time(e1):7.3f} ms | host enqueue {1e3 * (h1 - h0):7.3f} ms", flush=True)
kdist._MARKS = None
xchg.close()
dist.destroy_process_group()
