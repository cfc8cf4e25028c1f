"""Host-side logic of the multi-GPU paths (one process per GPU, torch.distributed for the plumbing).

* list-sharded IVFFlat scan: list ``l`` lives on rank ``l % world``; every rank keeps the
  GLOBAL list numbering with empty lists for the ones it does not own, so probe selection is identical
  on every rank; the only exchange is one all-gather of k (distance, id) pairs per rank + a k-way merge.
* sharded IVFFlat build: rows / samples are split contiguously; the library's all-reduce hook sums the
  per-rank centre sums, counts and change counters.

Tensor arguments may live on CPU (gloo tests) or CUDA (NCCL); nothing here computes distances.
"""
from __future__ import annotations

import numpy as np


def owner_of_list(l: int, world: int) -> int:
    return l % world


def shard_lists(list_offsets: np.ndarray, rank: int, world: int):
    """-> (row_mask [n] bool, local_offsets [lists+1]) keeping global list numbers (non-owned lists empty)."""
    off = np.asarray(list_offsets, dtype=np.int64)
    lists = off.shape[0] - 1
    lens = np.diff(off)
    own = (np.arange(lists) % world) == rank
    mask = np.zeros(int(off[-1]), dtype=bool)
    for l in np.nonzero(own)[0]:
        mask[off[l]:off[l + 1]] = True
    local = np.zeros(lists + 1, dtype=np.int64)
    local[1:] = np.cumsum(np.where(own, lens, 0))
    return mask, local


def shard_rows(n: int, rank: int, world: int):
    """contiguous split of n rows -> (lo, hi) of this rank"""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def merge_topk(dist_parts, ids_parts, k: int):
    """k-way merge of per-rank results.  dist_parts/ids_parts: tensors [world, nq, k] (missing = +inf / -1).
    Ties: lower rank first, then position (deterministic)."""
    import torch
    world, nq, kk = dist_parts.shape
    d = dist_parts.permute(1, 0, 2).reshape(nq, world * kk)
    i = ids_parts.permute(1, 0, 2).reshape(nq, world * kk)
    # stable sort keeps (rank, position) order among equal distances
    order = torch.sort(d, dim=1, stable=True).indices[:, :k]
    return torch.gather(d, 1, order), torch.gather(i, 1, order)


def all_gather_merge(dist_local, ids_local, k: int, group=None):
    """the one exchange of the list-sharded scan"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    gd = [torch.empty_like(dist_local) for _ in range(world)]
    gi = [torch.empty_like(ids_local) for _ in range(world)]
    dist.all_gather(gd, dist_local.contiguous(), group=group)
    dist.all_gather(gi, ids_local.contiguous(), group=group)
    return merge_topk(torch.stack(gd), torch.stack(gi), k)


def torch_allreduce_hook(device, group=None):
    """vb_allreduce_fn implemented with torch.distributed: wraps the raw device pointer in a tensor view."""
    import ctypes

    import torch
    import torch.distributed as dist

    np_dtypes = {0: (torch.float32, 4), 1: (torch.int32, 4), 2: (torch.int64, 8)}

    def hook(ptr, count, dtype_code):
        tdtype, size = np_dtypes[dtype_code]
        if device.type == "cuda":
            # zero-copy view of the library's device buffer through the CUDA array interface
            class _Buf:
                pass
            b = _Buf()
            b.__cuda_array_interface__ = {"shape": (int(count),), "typestr": {torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8"}[tdtype],
                                          "data": (int(ptr), False), "version": 2}
            t = torch.as_tensor(b, device=device)
            dist.all_reduce(t, group=group)
            # the library resumes on its own stream as soon as the hook returns
            torch.cuda.current_stream(device).synchronize()
        else:  # CPU tests: ptr is a host pointer
            buf = (ctypes.c_char * (int(count) * size)).from_address(int(ptr))
            t = torch.frombuffer(buf, dtype=tdtype)
            dist.all_reduce(t, group=group)
    return hook
