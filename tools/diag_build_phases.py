#!/usr/bin/env python3
"""Config-D sample phase in isolation (204 800 x 1536 samples, 4096 centres): a few k-means++ rounds and a few Lloyd
iterations, for a per-kernel launch list (run under `ncu --metrics gpu__time_duration.sum`) and wall-clock phases."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import pgvector_b200 as pv
    pv.init(0)
    dev = torch.device("cuda", 0)
    k, dim, ns = int(os.environ.get("K", 4096)), 1536, int(os.environ.get("NS", 204800))
    rounds = int(os.environ.get("ROUNDS", k))
    g = torch.Generator(device=dev).manual_seed(3)
    comp = torch.randn((k, dim), generator=g, device=dev)
    samp = torch.empty((ns, dim), device=dev)
    for lo in range(0, ns, 65536):
        hi = min(ns, lo + 65536)
        which = torch.randint(0, k, (hi - lo,), generator=g, device=dev)
        samp[lo:hi] = comp[which] + 0.3 * torch.randn((hi - lo, dim), generator=g, device=dev)
    torch.cuda.synchronize()
    t = pv.Table(pv.VECTOR, dim).append(samp)
    pv.synchronize()
    out = {}
    for flt in (1, 0):
        pv.set_option("pp_filter", flt)
        t0 = time.perf_counter()
        init = pv.kmeans_pp_init(t, pv.L2, rounds, seed=42)
        pv.synchronize()
        out[f"seeding_{rounds}_rounds_filter_{flt}_s"] = time.perf_counter() - t0
        out[f"pp_stats_filter_{flt}"] = pv.kmeans_pp_stats()
    pv.set_option("pp_filter", 1)
    if rounds < k:
        import numpy as np
        init = np.concatenate([init, samp[:k - rounds].cpu().numpy()])
    for it in (1, 3):
        t0 = time.perf_counter()
        c, iters = pv.kmeans(t, pv.L2, init, max_iter=it)
        pv.synchronize()
        out[f"lloyd_max_iter_{it}_s"] = time.perf_counter() - t0
        out[f"lloyd_max_iter_{it}_iters"] = iters
    print(json.dumps(out))


if __name__ == "__main__":
    main()
