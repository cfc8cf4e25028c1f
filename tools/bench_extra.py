#!/usr/bin/env python3
"""Secondary measurements (one JSON line each) for the rows of SURVEY section 8 that bench.py's headline
does not cover: k-means assign on the tensor cores (config D shape, one GPU's share), HNSW search
(configs C / E at reduced row counts: the graph is built by the CPU oracle, which bounds n), exact scan.

    python tools/bench_extra.py assign   [--rows N --dim D --k K]
    python tools/bench_extra.py hnsw     [--elem halfvec|bit --rows N --dim D --ef EF]
    python tools/bench_extra.py exact    [--rows N --dim D]

All timing with CUDA events on the library stream; inputs resident in HBM.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


def timed(pv, torch, stream, fn, warmup=2, steps=5):
    for _ in range(warmup):
        fn()
    pv.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    pv.synchronize()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def bench_assign(args):
    import torch
    import pgvector_b200 as pv
    pv.init(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(pv.stream_handle(), device=dev)
    g = torch.Generator(device=dev).manual_seed(5)
    comp = torch.randn((args.k, args.dim), generator=g, device=dev)
    which = torch.randint(0, args.k, (args.rows,), generator=g, device=dev)
    rows = comp[which] + 0.3 * torch.randn((args.rows, args.dim), generator=g, device=dev)
    centers = (comp + 0.05 * torch.randn((args.k, args.dim), generator=g, device=dev)).contiguous()
    torch.cuda.synchronize()
    t = pv.Table(pv.VECTOR, args.dim).append(rows)
    out = torch.empty(args.rows, dtype=torch.int32, device=dev)
    res = {}
    for name, tc in (("tcgen05_split_bf16", True), ("exact_fp32_cuda_cores", False)):
        pv.set_tensor_cores(tc)
        n_eff = args.rows if tc else min(args.rows, 131072)
        tt = t if tc else pv.Table(pv.VECTOR, args.dim).append(rows[:n_eff].contiguous())
        o = out[:n_eff]
        ms = timed(pv, torch, stream, lambda: pv._lib.check(pv.load().vb_assign_dev(tt.h, pv.L2_SQUARED, pv._ptr(centers), args.k, pv._ptr(o))),
                   warmup=1, steps=3)
        flops = 2.0 * n_eff * args.k * args.dim
        pv.prof_enable(True)
        pv.prof_read(pv.PROF_ASSIGN)
        pv._lib.check(pv.load().vb_assign_dev(tt.h, pv.L2_SQUARED, pv._ptr(centers), args.k, pv._ptr(o)))
        dev_ms, _ = pv.prof_read(pv.PROF_ASSIGN)     # pack + GEMM + re-check on the device, without host-side allocation
        pv.prof_enable(False)
        res[name] = {"ms": ms, "device_ms": dev_ms, "rows": n_eff, "useful_tflops": flops / ms / 1e9,
                     "useful_tflops_device": flops / dev_ms / 1e9, "rechecked_rows": pv.last_assign_rechecked()}
        if tc:
            a_tc = out.clone()
    pv.set_tensor_cores(True)
    # agreement between the two paths on the exact sample
    pv.set_tensor_cores(False)
    n_eff = res["exact_fp32_cuda_cores"]["rows"]
    ex = torch.empty(n_eff, dtype=torch.int32, device=dev)
    tt = pv.Table(pv.VECTOR, args.dim).append(rows[:n_eff].contiguous())
    pv._lib.check(pv.load().vb_assign_dev(tt.h, pv.L2_SQUARED, pv._ptr(centers), args.k, pv._ptr(ex)))
    pv.set_tensor_cores(True)
    agree = float((ex == a_tc[:n_eff]).float().mean().item())
    hbm, bf16_burst, bf16_sus, src = peaks()
    tc_ms = res["tcgen05_split_bf16"]["ms"]
    issued = 3 * 2.0 * args.rows * args.k * args.dim / tc_ms / 1e9      # three bf16 MMAs per fp32-accurate product
    print(json.dumps({"bench": "assign", "workload": f"assign {args.rows}x{args.dim} fp32 rows to {args.k} centres (L2)", "results": res,
                      "tc_vs_exact_agreement": agree,
                      "roofline": {"bound": "tensor", "achieved": issued, "peak": bf16_sus, "unit": "TFLOP/s", "frac": issued / bf16_sus,
                                   "peak_source": src + " bf16_tflops_sustained", "note": "issued bf16 MMA flops (3 per useful fp32-accurate MAC pair)"}}))


def bench_hnsw(args):
    import torch
    import oracle as O
    import pgvector_b200 as pv
    from tests.util import f32_to_half_bits, mixture
    pv.init(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(pv.stream_handle(), device=dev)
    # low intrinsic dimension (like bench.py): a graph index on it has a meaningful recall
    rng = np.random.default_rng(6)
    frame = np.linalg.qr(rng.standard_normal((args.dim, 16)))[0].astype(np.float32)
    x = (rng.standard_normal((args.rows, 16)).astype(np.float32) @ frame.T + 0.02 * rng.standard_normal((args.rows, args.dim)).astype(np.float32))
    q = (rng.standard_normal((args.queries, 16)).astype(np.float32) @ frame.T + 0.02 * rng.standard_normal((args.queries, args.dim)).astype(np.float32))
    if args.elem == "halfvec":
        elem, opclass, metric = O.HALFVEC, "halfvec_cosine_ops", O.NEG_IP
        rows = O.l2_normalize(O.HALFVEC, f32_to_half_bits(x))
        queries = O.l2_normalize(O.HALFVEC, f32_to_half_bits(q))
        rb = args.dim * 2
    else:
        elem, opclass, metric = O.BIT, "bit_hamming_ops", O.HAMMING
        rows, queries = O.binary_quantize(O.VECTOR, x), O.binary_quantize(O.VECTOR, q)
        rb = (args.dim + 7) // 8
    t0 = time.perf_counter()
    og = O.Hnsw(elem, metric, rows, m=16, ef_construction=64, seed=1, dim=args.dim)
    build_s = time.perf_counter() - t0
    g = og.export()
    gi = pv.HnswIndex(opclass, args.dim, m=16).load(rows[g["elem_row"]], g["levels"], g["nbr0"], g["upper_off"], g["upper"], g["entry"])
    qd = torch.from_numpy(queries).to(dev)
    k = 10
    ids = torch.empty((args.queries, k), dtype=torch.int64, device=dev)
    dist = torch.empty((args.queries, k), dtype=torch.float32, device=dev)
    nd = torch.empty(args.queries, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    ms = timed(pv, torch, stream, lambda: gi.search_into(qd, k, args.ef, ids, dist, nd), warmup=2, steps=5)
    ndist = float(nd.double().mean().item())
    # parity + recall on a sample
    ns = min(256, args.queries)
    wi, wd, wnd = og.search_batch(queries[:ns], args.ef, k, ties=O.TIES_TOTAL, threads=os.cpu_count() or 1)
    same = float(np.all(ids[:ns].cpu().numpy() == wi, axis=1).mean())
    erows = rows[g["elem_row"]]
    truth = [O.exact_topk(elem, metric, qq, erows, k, dim=args.dim)[0] for qq in queries[:64]]
    rec = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(ids[:64].cpu().numpy(), truth)) / (64 * k)
    t0 = time.perf_counter()
    og.search_batch(queries[:ns], args.ef, k, ties=O.TIES_PG, threads=os.cpu_count() or 1)
    cpu_qps = ns / (time.perf_counter() - t0)
    hbm, _, _, src = peaks()
    qps = args.queries / (ms / 1000.0)
    gbs = qps * (ndist * rb + (ndist / 16.0) * 32 * 4) / 1e9
    print(json.dumps({"bench": "hnsw", "workload": f"HNSW {opclass} {args.rows}x{args.dim}, m=16, ef_search={args.ef}, k={k} "
                                                   f"(graph built by the CPU oracle in {build_s:.0f}s)",
                      "queries_per_s": qps, "ms_per_batch": ms, "batch": args.queries, "distance_evals_per_query": ndist,
                      "recall_at_10": rec, "queries_identical_to_oracle": same,
                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "peak_source": src,
                                   "note": "latency-bound random gathers; bytes = n_dist*row + n_expand*lm*4"},
                      "cpu_baseline": {"value": cpu_qps, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port"}}))


def bench_ivf(args):
    """IVFFlat scan throughput for halfvec / bit rows (bench.py covers vector)."""
    import torch
    import pgvector_b200 as pv
    pv.init(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(pv.stream_handle(), device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    frame = torch.linalg.qr(torch.randn((args.dim, 16), generator=g, device=dev))[0]
    x = torch.randn((args.rows, 16), generator=g, device=dev) @ frame.T + 0.02 * torch.randn((args.rows, args.dim), generator=g, device=dev)
    q = torch.randn((args.queries, 16), generator=g, device=dev) @ frame.T + 0.02 * torch.randn((args.queries, args.dim), generator=g, device=dev)
    if args.elem == "halfvec":
        elem, opclass, kmetric, pmetric, rb = pv.HALFVEC, "halfvec_l2_ops", pv.L2, pv.L2_SQUARED, args.dim * 2
        rows_t, q_t = x.half().contiguous(), q.half().contiguous()
        rows_np = rows_t.cpu().numpy().view(np.uint16)
    else:
        elem, opclass, kmetric, pmetric, rb = pv.BIT, "bit_hamming_ops", pv.HAMMING, pv.HAMMING, args.dim // 8
        def pack(t):
            b = (t > 0).to(torch.uint8).reshape(t.shape[0], -1, 8)
            w = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.uint8, device=dev)
            return (b * w).sum(-1).to(torch.uint8).contiguous()
        rows_t, q_t = pack(x), pack(q)
        rows_np = rows_t.cpu().numpy()
    del x
    torch.cuda.synchronize()
    ns = min(args.rows, 50 * args.lists)
    ts = pv.Table(elem, args.dim).append(rows_np[:ns])
    init = pv.kmeans_pp_init(ts, kmetric, args.lists, seed=42)
    centers, iters = pv.kmeans(ts, kmetric, init, max_iter=100)
    ta = pv.Table(elem, args.dim).append(rows_t)
    assign = pv.assign(ta, pmetric, centers)
    ta.free()
    a = torch.from_numpy(assign).to(dev).long()
    order = torch.argsort(a, stable=True)
    counts = torch.bincount(a, minlength=args.lists).cpu()
    offsets = np.zeros(args.lists + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(counts.numpy())
    grouped = rows_t[order].contiguous()
    centers_t = torch.from_numpy(centers).to(dev)
    torch.cuda.synchronize()
    ix = pv.IvfflatIndex(opclass, args.dim, args.lists).load(centers_t, offsets, grouped, order.contiguous())
    k, B = 10, args.queries
    ids = torch.empty((B, k), dtype=torch.int64, device=dev)
    dist = torch.empty((B, k), dtype=torch.float32, device=dev)
    pv.prof_enable(True)
    pv.prof_read(pv.PROF_SCAN_ITEMS)
    ms = timed(pv, torch, stream, lambda: ix.search_into(q_t, k, args.probes, ids, dist), warmup=3, steps=10)
    scan_ms, scan_n = pv.prof_read(pv.PROF_SCAN_ITEMS)
    pv.prof_enable(False)
    cand = ix.last_candidates()
    hbm, _, _, src = peaks()
    gbs = cand * rb / (scan_ms / scan_n / 1000.0) / 1e9
    print(json.dumps({"bench": "ivf", "workload": f"IVFFlat {opclass} {args.rows}x{args.dim}, lists={args.lists}, probes={args.probes}, k={k}, {B} queries per batch "
                                                  f"(k-means {iters} it; list sizes {int(counts.min())}/{int(counts.float().mean())}/{int(counts.max())})",
                      "queries_per_s": B / (ms / 1000.0), "ms_per_batch": ms, "candidates_per_query": cand / B,
                      "roofline": {"bound": "hbm", "kernel": "list scan", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                                   "bytes_per_launch": cand * rb, "avg_launch_ms": scan_ms / scan_n, "share_of_step": scan_ms / scan_n / ms, "peak_source": src}}))


def bench_kmeans(args):
    """k-means of config D on one GPU: 50 * lists samples x dim, lists centres (k-means++ seeding + Lloyd iterations)."""
    import torch
    import pgvector_b200 as pv
    pv.init(0)
    dev = torch.device("cuda", 0)
    n = 50 * args.k
    g = torch.Generator(device=dev).manual_seed(5)
    frame = torch.linalg.qr(torch.randn((args.dim, 16), generator=g, device=dev))[0]
    rows = torch.randn((n, 16), generator=g, device=dev) @ frame.T + 0.02 * torch.randn((n, args.dim), generator=g, device=dev)
    torch.cuda.synchronize()
    t = pv.Table(pv.VECTOR, args.dim).append(rows)
    pv.synchronize()
    # one untimed pass of each entry point first: workspace growth, pinned staging and lazy module loading are
    # one-time costs of the process, not of a build
    pv.kmeans(t, pv.L2, pv.kmeans_pp_init(t, pv.L2, min(args.k, 64), seed=1), max_iter=1)
    pv.kmeans(t, pv.L2, rows[:args.k].cpu().numpy(), max_iter=1)
    t0 = time.perf_counter()
    init = pv.kmeans_pp_init(t, pv.L2, args.k, seed=42)
    pp_s = time.perf_counter() - t0
    pv.prof_enable(True)
    pv.prof_read(pv.PROF_ASSIGN)
    t0 = time.perf_counter()
    centers, iters = pv.kmeans(t, pv.L2, init, max_iter=args.iters)
    km_s = time.perf_counter() - t0
    assign_ms, assign_n = pv.prof_read(pv.PROF_ASSIGN)
    pv.prof_enable(False)
    _, _, bf16_sus, src = peaks()
    flops = 2.0 * n * args.k * args.dim
    issued = 3 * flops / (assign_ms / assign_n) / 1e9
    print(json.dumps({"bench": "kmeans", "workload": f"k-means {n}x{args.dim} fp32 samples, {args.k} centres (config D sample phase on one GPU)",
                      "kmeanspp_s": pp_s, "kmeans_s": km_s, "iterations": iters, "s_per_iteration": km_s / max(iters, 1),
                      "assign_ms_per_iteration": assign_ms / assign_n, "rechecked_rows_last": pv.last_assign_rechecked(),
                      "roofline": {"bound": "tensor", "kernel": "assign step (pack + tcgen05 GEMM + re-check)", "achieved": issued, "peak": bf16_sus,
                                   "unit": "TFLOP/s", "frac": issued / bf16_sus, "peak_source": src + " bf16_tflops_sustained"}}))


def bench_exact(args):
    import torch
    import pgvector_b200 as pv
    pv.init(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(pv.stream_handle(), device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    rows = torch.randn((args.rows, args.dim), generator=g, device=dev)
    q = torch.randn((args.queries, args.dim), generator=g, device=dev)
    torch.cuda.synchronize()
    t = pv.Table(pv.VECTOR, args.dim).append(rows)
    k = 10
    ids = torch.empty((args.queries, k), dtype=torch.int64, device=dev)
    dist = torch.empty((args.queries, k), dtype=torch.float32, device=dev)
    ms = timed(pv, torch, stream, lambda: pv._lib.check(pv.load().vb_exact_topk_dev(t.h, pv.L2, pv._ptr(q), args.queries, k, pv._ptr(ids), pv._ptr(dist))))
    hbm, _, _, src = peaks()
    gbs = args.queries * args.rows * args.dim * 4 / (ms / 1000.0) / 1e9
    print(json.dumps({"bench": "exact", "workload": f"exact L2 top-{k} over {args.rows}x{args.dim} fp32, {args.queries} queries per batch",
                      "queries_per_s": args.queries / (ms / 1000.0), "ms_per_batch": ms,
                      "roofline": {"bound": "hbm (L2-resident when the table fits in 126 MB)", "achieved": gbs, "peak": hbm, "unit": "GB/s",
                                   "frac": gbs / hbm, "peak_source": src}}))


def bench_sparse(args):
    """sparsevec exact scan: resident CSR table, a batch of sparse queries per call (host buffers for the queries and the
    results); the roofline is the table's CSR bytes read once per query (8 bytes per stored entry)"""
    import torch
    import pgvector_b200 as pv
    import oracle as O
    S = pv.sparsevec
    pv.init(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(pv.stream_handle(), device=dev)
    rng = np.random.default_rng(1)
    n, dim, row_nnz, q_nnz, nq, k = args.rows, args.dim, args.nnz, 2 * args.nnz, args.queries, 10

    def ascending(count, nnz):
        """count rows of nnz strictly ascending indices in [0, dim): cumulative sums of gaps >= 1"""
        gaps = rng.integers(1, max(2, dim // nnz), size=(count, nnz))
        return (np.cumsum(gaps, axis=1) - 1).astype(np.int32)

    off = np.arange(n + 1, dtype=np.int64) * row_nnz
    idx = np.empty(n * row_nnz, dtype=np.int32)
    for r0 in range(0, n, 100_000):
        r1 = min(n, r0 + 100_000)
        idx[off[r0]:off[r1]] = ascending(r1 - r0, row_nnz).ravel()
    val = rng.standard_normal(off[-1]).astype(np.float32)
    table = S.SparseTable(dim).append(S.SparseRows(dim, off, idx, val))
    qidx = ascending(nq, q_nnz)
    qs = [S.SparseVector(dim, qidx[i], rng.standard_normal(q_nnz).astype(np.float32)) for i in range(nq)]
    Q = S.SparseRows.from_vectors(qs, dim)
    out = {}
    for name, metric in (("l2", O.L2), ("ip", O.NEG_IP)):
        ms = timed(pv, torch, stream, lambda: table.exact_topk(metric, Q, k), warmup=2, steps=5)
        out[name] = ms
    ids, dist = table.exact_topk(O.L2, Q, k)
    # parity on a few queries against the oracle
    same = 0
    for qi in range(min(8, nq)):
        d = O.sparse_distance_batch(O.L2, (qs[qi].indices, qs[qi].values), off, idx, val)
        want = np.argsort(d, kind="stable")[:k]
        same += int(np.array_equal(ids[qi], want))
    hbm, _, _, src = peaks()
    csr_bytes = off[-1] * 8 + (n + 1) * 8
    line = {"bench": "sparse", "workload": f"sparsevec exact top-{k}: {n} rows x {row_nnz} stored entries (dim {dim}), {nq} queries of ~{q_nnz} entries per call",
            "queries_per_s": {m: nq / (ms / 1000.0) for m, ms in out.items()}, "ms_per_call": out,
            "parity": {"queries_checked": min(8, nq), "identical_top_k_ids": same},
            "roofline": {"bound": "hbm", "kernel": "sparse_scan_kernel + segment_topk_kernel", "unit": "GB/s", "peak": hbm, "peak_source": src,
                         "traffic_per_query": int(csr_bytes), "achieved": {m: nq * csr_bytes / (ms / 1000.0) / 1e9 for m, ms in out.items()},
                         "frac": {m: nq * csr_bytes / (ms / 1000.0) / 1e9 / hbm for m, ms in out.items()},
                         "note": "the call also uploads the queries, writes and re-reads the nq x rows key matrix for the selection and returns the results"}}
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["assign", "hnsw", "exact", "ivf", "kmeans", "sparse"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--lists", type=int, default=1000)
    ap.add_argument("--probes", type=int, default=10)
    ap.add_argument("--rows", type=int, default=None)
    ap.add_argument("--dim", type=int, default=None)
    ap.add_argument("--k", type=int, default=4096)
    ap.add_argument("--elem", default="halfvec")
    ap.add_argument("--ef", type=int, default=100)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--nnz", type=int, default=100)
    a = ap.parse_args()
    if a.what == "sparse":
        a.rows = a.rows or 1_000_000
        a.dim = a.dim or 100_000
        a.queries = min(a.queries, 64)
        bench_sparse(a)
    elif a.what == "assign":
        a.rows = a.rows or 1_250_000          # one GPU's share of config D (10M rows / 8)
        a.dim = a.dim or 1536
        bench_assign(a)
    elif a.what == "hnsw":
        a.rows = a.rows or 100_000
        a.dim = a.dim or (768 if a.elem == "halfvec" else 1024)
        bench_hnsw(a)
    elif a.what == "kmeans":
        a.dim = a.dim or 1536
        bench_kmeans(a)
    elif a.what == "ivf":
        a.rows = a.rows or 1_000_000
        a.dim = a.dim or (1536 if a.elem == "halfvec" else 1024)
        bench_ivf(a)
    else:
        a.rows = a.rows or 10_000
        a.dim = a.dim or 128
        bench_exact(a)
