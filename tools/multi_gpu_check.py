#!/usr/bin/env python3
"""Multi-GPU correctness check (run under torchrun, NCCL):
   1. list-sharded IVFFlat scan + all-gather merge == the single-GPU scan of the same index;
   2. row-sharded k-means with the NCCL all-reduce hook == single-GPU k-means from the same initial centres.
Prints one JSON line on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tools/multi_gpu_check.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import pgvector_b200 as pv
    from pgvector_b200 import sharding

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pv.init(local)

    rng = np.random.default_rng(3)          # same data on every rank
    k, dim, n, nq = 64, 256, 60000, 512
    comp = rng.standard_normal((8 * k, dim)).astype(np.float32)
    rows = (comp[rng.integers(0, 8 * k, n)] + 0.3 * rng.standard_normal((n, dim))).astype(np.float32)
    queries = (comp[rng.integers(0, 8 * k, nq)] + 0.3 * rng.standard_normal((nq, dim))).astype(np.float32)

    # ---- k-means: single GPU vs row-sharded with the all-reduce hook
    samples = rows[:20000]
    t_all = pv.Table(pv.VECTOR, dim).append(samples)
    init = pv.kmeans_pp_init(t_all, pv.L2, k, seed=7)
    c_single, it_single = pv.kmeans(t_all, pv.L2, init, max_iter=30)
    lo, hi = sharding.shard_rows(len(samples), rank, world)
    t_part = pv.Table(pv.VECTOR, dim).append(samples[lo:hi])
    hook = sharding.torch_allreduce_hook(dev)
    c_shard, it_shard = pv.kmeans(t_part, pv.L2, init, max_iter=30, allreduce=hook)
    km_max_diff = float(np.abs(c_single - c_shard).max())

    # ---- scan: full index vs list shards + merge
    assign = pv.assign(pv.Table(pv.VECTOR, dim).append(rows), pv.L2_SQUARED, c_single)
    order = np.argsort(assign, kind="stable")
    counts = np.bincount(assign, minlength=k)
    offsets = np.zeros(k + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(counts)
    grouped, ids = rows[order], order.astype(np.int64)
    full = pv.IvfflatIndex("vector_l2_ops", dim, k).load(c_single, offsets, grouped, ids)
    f_ids, f_dist = full.search(queries, k=10, probes=8)
    mask, local_off = sharding.shard_lists(offsets, rank, world)
    part = pv.IvfflatIndex("vector_l2_ops", dim, k).load(c_single, local_off, grouped[mask], ids[mask])
    p_ids, p_dist = part.search(queries, k=10, probes=8)
    md, mi = sharding.all_gather_merge(torch.from_numpy(p_dist).to(dev), torch.from_numpy(p_ids).to(dev), 10)
    scan_same = bool(np.array_equal(mi.cpu().numpy(), f_ids) and np.allclose(md.cpu().numpy(), f_dist, rtol=1e-6))

    if rank == 0:
        print(json.dumps({"world": world, "kmeans_iters_single": it_single, "kmeans_iters_sharded": it_shard,
                          "kmeans_centre_max_abs_diff": km_max_diff, "sharded_scan_equals_single": scan_same,
                          "ok": bool(scan_same and km_max_diff < 1e-3 and it_single == it_shard)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
