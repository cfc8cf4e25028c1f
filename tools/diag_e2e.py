#!/usr/bin/env python3
"""Where the end-to-end step of config B goes (host-buffer C ABI): H2D alone, the search alone, pipelined and plain host calls,
each with wall-clock and CUDA-event time per step.  Run on the GPU box: python tools/diag_e2e.py [--rows N]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=60)
    a = ap.parse_args()
    sys.argv = [sys.argv[0], "--no-cpu", "--no-recall", "--no-extras", "--law", "rank16"]
    args = bench.parse_args()
    env = bench.Env(args)
    torch, pv = env.torch, env.pv
    rows, queries = bench.make_dataset(args, "rank16", env.dev)
    centers, offsets, grouped, order, how = bench.build_index_arrays(args, "rank16", rows, pv)
    del rows
    B, k = min(args.batch, args.queries), args.k
    ix = pv.IvfflatIndex("vector_l2_ops", args.dim, args.lists).load(centers, offsets, grouped, order)
    nb = max(1, args.queries // B)
    qb = [queries[i * B:(i + 1) * B].contiguous() for i in range(nb)]
    ids_dev = torch.empty((B, k), dtype=torch.int64, device=env.dev)
    dist_dev = torch.empty((B, k), dtype=torch.float32, device=env.dev)
    q_host = [torch.empty((B, args.dim), dtype=torch.float32).pin_memory().copy_(x.cpu()) for x in qb[:4]]
    q_np = [t.numpy() for t in q_host]
    q_pageable = [np.array(x) for x in q_np]
    ids_h = torch.empty((B, k), dtype=torch.int64).pin_memory().numpy()
    dist_h = torch.empty((B, k), dtype=torch.float64).pin_memory().numpy()
    out = {"pinned": [bool(t.is_pinned()) for t in q_host]}

    def timed(name, fn, steps=a.steps, warm=5):
        for i in range(warm):
            fn(i)
        pv.synchronize()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(env.stream)
        for i in range(steps):
            fn(warm + i)
        e1.record(env.stream)
        pv.synchronize()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1000 / steps
        out[name] = {"wall_ms_per_step": wall, "event_ms_per_step": e0.elapsed_time(e1) / steps}

    # 1. H2D of one query batch alone (torch copy of the pinned tensor, then the library's prefetch + wait)
    dst = torch.empty((B, args.dim), dtype=torch.float32, device=env.dev)

    def h2d(i):
        dst.copy_(q_host[i % 4], non_blocking=True)
        torch.cuda.synchronize()
    timed("h2d_torch_pinned_sync", h2d)

    def pre_only(i):
        ix.prefetch_queries(q_np[i % 4], i % 2)
        pv.synchronize()
    timed("prefetch_only (stream sync does not wait for the copy stream)", pre_only)
    # 2. device-resident search
    timed("search_dev", lambda i: ix.search_into(qb[i % nb], k, args.probes, ids_dev, dist_dev))
    # 3. prefetched search without a concurrent copy: prefetch, wait for it, then search
    def pre_then_search(i):
        ix.prefetch_queries(q_np[i % 4], i % 2)
        torch.cuda.synchronize()
        t = time.perf_counter()
        ix.search_prefetched_into(i % 2, k, args.probes, ids_h, dist_h)
        pre_then_search.acc += time.perf_counter() - t
    pre_then_search.acc = 0.0
    timed("prefetch_wait_then_search", pre_then_search)
    out["prefetch_wait_then_search"]["search_call_wall_ms"] = pre_then_search.acc * 1000 / (a.steps + 5)
    # 4. the bench's pipelined step
    n = [0]
    ix.prefetch_queries(q_np[0], 0)

    def piped(_):
        i = n[0]
        n[0] += 1
        t = time.perf_counter()
        ix.prefetch_queries(q_np[(i + 1) % 4], (i + 1) % 2)
        piped.pre += time.perf_counter() - t
        ix.search_prefetched_into(i % 2, k, args.probes, ids_h, dist_h)
    piped.pre = 0.0
    timed("pipelined", piped)
    out["pipelined"]["prefetch_call_wall_ms"] = piped.pre * 1000 / (a.steps + 5)
    # 5. plain host call, pinned and pageable queries
    timed("search_host_pinned", lambda i: ix.search_host_into(q_np[i % 4], k, args.probes, ids_h, dist_h))
    timed("search_host_pageable", lambda i: ix.search_host_into(q_pageable[i % 4], k, args.probes, ids_h, dist_h))
    # 6. the same loops with bench.py's clock sampler (nvidia-smi -lms) running beside them, and with the profiling brackets on
    sampler = bench.ClockSampler(env.local)
    sampler.start()
    time.sleep(0.5)
    n[0] = 0
    ix.prefetch_queries(q_np[0], 0)
    timed("pipelined_with_clock_sampler", piped)
    timed("search_dev_with_clock_sampler", lambda i: ix.search_into(qb[i % nb], k, args.probes, ids_dev, dist_dev))
    out["sampler"] = sampler.stop()
    time.sleep(0.3)
    n[0] = 0
    ix.prefetch_queries(q_np[0], 0)
    timed("pipelined_after_sampler_stopped", piped)
    pv.prof_enable(True)
    timed("search_dev_with_prof_brackets", lambda i: ix.search_into(qb[i % nb], k, args.probes, ids_dev, dist_dev))
    pv.tc_traffic(True, read=True)
    timed("search_dev_with_prof_brackets_and_traffic_accounting", lambda i: ix.search_into(qb[i % nb], k, args.probes, ids_dev, dist_dev))
    pv.tc_traffic(False, read=True)
    pv.prof_enable(False)
    # 7. bench-like order: 700 device steps first, then the pipelined loop with 100 steps
    for i in range(700):
        ix.search_into(qb[i % nb], k, args.probes, ids_dev, dist_dev)
    n[0] = 0
    ix.prefetch_queries(q_np[0], 0)
    timed("pipelined_after_700_device_steps", piped, steps=100, warm=3)
    print(json.dumps(out, indent=1))
    env.close()


if __name__ == "__main__":
    main()
