"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: list sharding + all-gather merge of the
IVFFlat scan, and the all-reduce semantics of the sharded k-means centre update.  The per-rank scanner
is the CPU oracle here (no GPU in this container); on the GPU the same host code drives libvecb200."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from pgvector_b200 import sharding
from tests.util import build_ivf_arrays, mixture


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2, port=29631):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _dataset():
    rows, centers = mixture(6000, 24, 16, seed=3)
    queries, _ = mixture(40, 24, 16, seed=4)
    assign = O.ivf_assign(O.VECTOR, O.L2_SQUARED, rows, centers)
    grouped, ids, offsets = build_ivf_arrays(rows, assign, 16)
    return centers, grouped, ids, offsets, queries


def _sharded_scan(rank, world):
    centers, grouped, ids, offsets, queries = _dataset()
    mask, local_off = sharding.shard_lists(offsets, rank, world)
    ix = O.Ivf(O.VECTOR, O.L2_SQUARED, centers, local_off, grouped[mask], ids[mask])
    li, ld = ix.search_batch(queries, probes=5, k=10)
    d, i = sharding.all_gather_merge(torch.from_numpy(ld), torch.from_numpy(li), 10)
    return d.numpy(), i.numpy()


def test_list_sharded_scan_equals_unsharded():
    centers, grouped, ids, offsets, queries = _dataset()
    full = O.Ivf(O.VECTOR, O.L2_SQUARED, centers, offsets, grouped, ids)
    wi, wd = full.search_batch(queries, probes=5, k=10)
    res = _run(_sharded_scan)
    for d, i in res:                      # every rank ends with the same merged answer
        assert np.array_equal(d, wd)
        assert np.array_equal(i, wi)


def test_shard_lists_partitions_rows():
    _, grouped, ids, offsets, _ = _dataset()
    masks = [sharding.shard_lists(offsets, r, 4)[0] for r in range(4)]
    assert np.array_equal(np.sum(masks, axis=0), np.ones(len(ids)))
    for r in range(4):
        m, lo = sharding.shard_lists(offsets, r, 4)
        assert lo[-1] == m.sum()
        lens = np.diff(lo)
        assert np.all(lens[np.arange(16) % 4 != r] == 0)


def _sharded_centre_update(rank, world):
    x, c = mixture(3000, 8, 6, seed=7)
    lo, hi = sharding.shard_rows(len(x), rank, world)
    a = O.ivf_assign(O.VECTOR, O.L2_SQUARED, x[lo:hi], c)
    sums = np.zeros((6, 8), np.float32)
    np.add.at(sums, a, x[lo:hi])
    counts = np.bincount(a, minlength=6).astype(np.int32)
    hook = sharding.torch_allreduce_hook(torch.device("cpu"))
    hook(sums.ctypes.data, sums.size, 0)      # the calls libvecb200 makes once per Lloyd iteration
    hook(counts.ctypes.data, counts.size, 1)
    return sums, counts


def test_allreduce_hook_sums_partial_centres():
    x, c = mixture(3000, 8, 6, seed=7)
    a = O.ivf_assign(O.VECTOR, O.L2_SQUARED, x, c)
    want = np.zeros((6, 8), np.float32)
    np.add.at(want, a, x)
    res = _run(_sharded_centre_update, port=29633)
    for sums, counts in res:
        assert np.array_equal(counts, np.bincount(a, minlength=6))
        assert np.allclose(sums, want, rtol=1e-5, atol=1e-4)


def test_merge_topk_ties_and_padding():
    d = torch.tensor([[[1.0, 3.0, float("inf")]], [[1.0, 2.0, 4.0]]])
    i = torch.tensor([[[10, 11, -1]], [[20, 21, 22]]])
    md, mi = sharding.merge_topk(d, i, 4)
    assert md.tolist() == [[1.0, 1.0, 2.0, 3.0]]
    assert mi.tolist() == [[10, 20, 21, 11]]
