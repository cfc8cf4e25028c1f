#!/usr/bin/env python3
"""Checks of the library's own NCCL paths, run under torchrun on >= 2 GPUs:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py

1. vb_comm all-reduce / all-gather;  2. k-means++ over row-sharded samples picks the rows the single-GPU seeding picks
from the same draws;  3. k-means over sharded samples == single-GPU k-means;  4. the list-sharded search returns the
single-GPU search's neighbours."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def splitmix_uniforms(seed, count):
    st = (seed ^ 0x5851f42d4c957f2d) & 0xFFFFFFFFFFFFFFFF
    out = []
    for _ in range(count):
        st = (st + 0x9e3779b97f4a7c15) & 0xFFFFFFFFFFFFFFFF
        z = st
        z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & 0xFFFFFFFFFFFFFFFF
        z ^= z >> 31
        out.append((z >> 11) * (1.0 / 9007199254740992.0))
    return out


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import pgvector_b200 as pv
    pv.init(local)
    ident = [pv.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    pv.comm_init(ident[0], rank, world)
    assert pv.comm_world() == world
    lib = pv.load()
    import ctypes as C
    ok = True

    def say(name, good, detail=""):
        nonlocal ok
        ok = ok and good
        if rank == 0:
            print(("PASS " if good else "FAIL ") + name + (" " + str(detail) if detail else ""), flush=True)

    # 1. collectives
    x = torch.full((1000,), float(rank + 1), device=dev)
    pv._after_torch(x)
    pv._lib.check(lib.vb_comm_allreduce(C.c_void_p(x.data_ptr()), 1000, 0))
    pv.synchronize()
    say("allreduce", bool((x == world * (world + 1) / 2).all().item()))
    send = torch.full((16,), rank, dtype=torch.int32, device=dev)
    recv = torch.empty((world * 16,), dtype=torch.int32, device=dev)
    pv._after_torch(send)
    pv._lib.check(lib.vb_comm_allgather(C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()), 64))
    pv.synchronize()
    say("allgather", bool((recv.view(world, 16) == torch.arange(world, device=dev, dtype=torch.int32)[:, None]).all().item()))

    # 2. k-means++ over sharded samples
    rng = np.random.default_rng(7)
    n, dim, k = 24000, 96, 64
    comps = rng.standard_normal((40, dim)).astype(np.float32)
    samples = (comps[rng.integers(0, 40, n)] + 0.3 * rng.standard_normal((n, dim))).astype(np.float32)
    cut = [n * r // world + (37 if 0 < r < world else 0) for r in range(world + 1)]      # uneven slices
    mine = samples[cut[rank]:cut[rank + 1]]
    t = pv.Table(pv.VECTOR, dim).append(mine)
    got = pv.kmeans_pp_init(t, pv.L2, k, seed=42)
    us = splitmix_uniforms(42, k)
    first = min(int(us[0] * n), n - 1)
    # the unsharded seeding needs a process without the communicator's row sharding: the explicit-draws entry point
    tf = pv.Table(pv.VECTOR, dim).append(samples)
    want, picked = pv.kmeans_pp_init_draws(tf, pv.L2, k, first, np.array(us[1:k]))
    same = np.all(got == want, axis=1)
    say("kmeans++ sharded == single", bool(same.all()), f"{int(same.sum())}/{k} centres equal")

    # 3. Lloyd over sharded samples
    c_sh, it_sh = pv.kmeans(t, pv.L2, want, max_iter=30)
    pv.comm_free()
    c_one, it_one = pv.kmeans(tf, pv.L2, want, max_iter=30)
    # partial sums are added in rank order, not in sample order: one borderline sample may change cluster, which moves
    # two centres by ~|x| / members; everything else must agree and the objective must be the same
    dc = np.max(np.abs(c_sh - c_one), axis=1)

    def objective(cent):
        d = (samples ** 2).sum(1)[:, None] - 2.0 * samples @ cent.T + (cent ** 2).sum(1)[None, :]
        return float(d.min(1).sum())
    o_sh, o_one = objective(c_sh.astype(np.float64)), objective(c_one.astype(np.float64))
    say("k-means sharded ~ single", float((dc < 1e-3).mean()) >= 0.9 and abs(o_sh - o_one) <= 1e-4 * o_one and abs(it_sh - it_one) <= 2,
        f"centres within 1e-3: {int((dc < 1e-3).sum())}/{k}, max |dc| {dc.max():.2e}, objective {o_sh:.6g} vs {o_one:.6g}, "
        f"iterations {it_sh} vs {it_one}")

    # 4. list-sharded search (a fresh communicator: the id was consumed)
    ident = [pv.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ident, src=0)
    pv.comm_init(ident[0], rank, world)
    lists, probes, kk = 200, 10, 10
    cen = samples[np.random.default_rng(3).choice(n, lists, replace=False)].copy()
    assign = pv.assign(tf, pv.L2_SQUARED, cen).astype(np.int64)
    order = np.argsort(assign, kind="stable")
    counts = np.bincount(assign, minlength=lists)
    off = np.zeros(lists + 1, dtype=np.int64)
    off[1:] = np.cumsum(counts)
    grouped = samples[order]
    full = pv.IvfflatIndex("vector_l2_ops", dim, lists).load(cen, off, grouped, order)
    keep = (np.arange(lists) % world) == rank
    sel = np.zeros(n, dtype=bool)
    for l in range(lists):
        if keep[l]:
            sel[off[l]:off[l + 1]] = True
    off_l = np.zeros(lists + 1, dtype=np.int64)
    off_l[1:] = np.cumsum(np.where(keep, counts, 0))
    shard = pv.IvfflatIndex("vector_l2_ops", dim, lists).load(cen, off_l, grouped[sel], order[sel])
    q = torch.from_numpy((comps[rng.integers(0, 40, 600)] + 0.3 * rng.standard_normal((600, dim))).astype(np.float32)).to(dev)
    torch.cuda.synchronize()
    for impl in (3, 4):
        pv.set_option("scan_impl", impl)
        wi, wd = full.search(q, k=kk, probes=probes)
        ids = torch.empty((600, kk), dtype=torch.int64, device=dev)
        dd = torch.empty((600, kk), dtype=torch.float32, device=dev)
        shard.search_sharded_into(q, kk, probes, ids, dd)
        pv.synchronize()
        agree = float((ids == wi).float().mean().item())
        derr = float(((dd - wd).abs() / wd.abs().clamp(min=1e-20)).max().item())
        say(f"sharded search == single (scan_impl {impl})", agree > 0.999 and derr < 1e-5, f"id agreement {agree:.5f}, max rel dist err {derr:.1e}")
        hi = np.empty((600, kk), dtype=np.int64)
        hd = np.empty((600, kk), dtype=np.float64)
        shard.search_sharded_host_into(q.cpu().numpy(), kk, probes, hi, hd)
        say(f"host-buffer sharded search == device variant (scan_impl {impl})", bool(np.array_equal(hi, ids.cpu().numpy())))
    pv.set_option("scan_impl", 2)
    # 5. exact scan over row-sharded tables == exact scan over the whole table
    lo, hi = n * rank // world, n * (rank + 1) // world
    tl = pv.Table(pv.VECTOR, dim).append(samples[lo:hi])
    e_ids, e_d = tl.exact_topk_sharded(pv.L2_SQUARED, q[:128].contiguous(), 10, lo)
    w_ids, w_d = tf.exact_topk(pv.L2_SQUARED, q[:128].contiguous(), 10)
    say("sharded exact scan == single", bool((e_ids == w_ids).float().mean().item() > 0.999 and torch.allclose(e_d, w_d, rtol=1e-6)))
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    pv.comm_free()
    dist.destroy_process_group()
    if rank == 0:
        print("ALL PASS" if flag.item() == 1.0 else "SOME FAILED", flush=True)
    return 0 if flag.item() == 1.0 else 1


if __name__ == "__main__":
    sys.exit(main())
