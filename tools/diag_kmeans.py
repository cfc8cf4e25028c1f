#!/usr/bin/env python3
"""Diagnostics for the GPU k-means / assign path on the bench data law (prints one JSON line per stage)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stats(a, k):
    c = np.bincount(a, minlength=k)
    return {"min": int(c.min()), "max": int(c.max()), "empty": int((c == 0).sum()), "p99": int(np.percentile(c, 99))}


def main():
    import torch
    import pgvector_b200 as pv
    pv.init(0)
    dev = torch.device("cuda", 0)
    k, dim, ns, n = int(os.environ.get("K", 1000)), 1536, 50000, int(os.environ.get("N", 300000))
    g = torch.Generator(device=dev).manual_seed(3)
    comp = torch.randn((k, dim), generator=g, device=dev)
    which = torch.randint(0, k, (n,), generator=g, device=dev)
    rows = comp[which] + 0.3 * torch.randn((n, dim), generator=g, device=dev)
    samp = rows[:ns].contiguous()
    torch.cuda.synchronize()
    t = pv.Table(pv.VECTOR, dim).append(samp)
    pv.synchronize()
    init = pv.kmeans_pp_init(t, pv.L2, k, seed=42)
    init_t = torch.from_numpy(init).to(dev)
    # which mixture component does each initial centre come from?
    owner = torch.cdist(init_t, comp).argmin(1).cpu().numpy()
    print(json.dumps({"stage": "kmeans++", "distinct_components": int(len(set(owner.tolist()))), "k": k}))
    out = {}
    for tc in (False, True):
        pv.set_tensor_cores(tc)
        a = pv.assign(t, pv.L2_SQUARED, init_t).cpu().numpy()
        out[tc] = a
        print(json.dumps({"stage": "assign samples to init", "tensor_cores": tc, "rechecked": pv.last_assign_rechecked(), **stats(a, k)}))
    print(json.dumps({"stage": "tc vs exact on samples", "agreement": float((out[True] == out[False]).mean())}))
    for tc in (False, True):
        pv.set_tensor_cores(tc)
        c, it = pv.kmeans(t, pv.L2, init, max_iter=50)
        ct = torch.from_numpy(c).to(dev)
        a = pv.assign(t, pv.L2_SQUARED, ct).cpu().numpy()
        print(json.dumps({"stage": "kmeans", "tensor_cores": tc, "iters": it, **stats(a, k)}))
        tr = pv.Table(pv.VECTOR, dim).append(rows)
        pv.set_tensor_cores(False)
        ae = pv.assign(tr, pv.L2_SQUARED, ct).cpu().numpy()
        pv.set_tensor_cores(True)
        at = pv.assign(tr, pv.L2_SQUARED, ct).cpu().numpy()
        print(json.dumps({"stage": "assign all rows", "centres_from_tc": tc, "exact": stats(ae, k), "tc": stats(at, k),
                          "agreement": float((ae == at).mean()), "rechecked": pv.last_assign_rechecked()}))
        bad = np.nonzero(ae != at)[0]
        if len(bad):
            print(json.dumps({"first_disagreeing_rows": bad[:20].tolist(), "slab_rows": 148 * 4 * 128}))
        tr.free()
    pv.set_tensor_cores(True)


if __name__ == "__main__":
    main()
