#!/usr/bin/env python3
"""bench.py -- IVFFlat scan throughput (BASELINE.json metric) on B200.

Workload (config B of BASELINE.json / SURVEY.md section 8d): 1,000,000 x 1536-d fp32 rows (seed 3; data law in
make_dataset: 1536-d vectors of intrinsic dimension 16, see DESIGN.md section 5 for why not the isotropic
mixture), ivfflat vector_l2_ops, lists = 1000 built by the library's own k-means++ / k-means / assign,
probes = 10, k = 10; 10,000 queries from the same law (seed 4).  A "step" is one batch of --batch queries
through the hot path (probe selection + list scan + top-k).

  value : queries/s with queries and outputs resident in HBM (device pointers, C ABI *_dev call)
  e2e   : queries/s through the host-buffer C ABI (vb_ivf_prefetch_queries for the next batch +
          vb_ivf_search_prefetched for the current one: every step has its H2D copy of 2048 queries from
          pinned memory and its D2H of ids + distances inside the timed region, the copy overlapping the
          previous batch's device work; the last step is compared with plain vb_ivf_search); the index
          image stays resident in HBM (uploaded once per index version, like shared_buffers; upload time
          in config.index_upload_s)
  roofline : SURVEY 8(d)'s algorithmic bytes of the list scan (one row read per distance, not amortised
          over the batch) / the kernel's CUDA-event time vs MEASURED_PEAKS.json, plus `dram`: the ncu
          DRAM bytes of the same launch / the same time (the batched kernels read each probed list once
          per batch, so `frac` > 1 is the reuse factor and `dram.frac` the physical roofline)
  cpu_baseline : the oracle port of the same scan on the host cores (bounded sample)

`--impl reference` times only the CPU arm (oracle port of src/ivfscan.c; PostgreSQL itself is
not installable here).  Under torchrun every rank shards the lists
(list l lives on rank l % N), all ranks see all queries, and per-query top-k lists are
exchanged with one NCCL all-gather and merged on the GPU ("strong" scaling: same index).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

try:    # the metric is BASELINE.json's, verbatim
    METRIC = json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
except Exception:
    METRIC = "IVFFlat 1M\u00d71536d queries/sec at 1/2/4/8 GPU; recall@10; HBM GB/s vs roofline"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=1536)
    ap.add_argument("--lists", type=int, default=1000)
    ap.add_argument("--latent-dim", type=int, default=16, help="intrinsic dimension of the default synthetic data law")
    ap.add_argument("--components", type=int, default=0,
                    help="> 0: use the Gaussian-mixture law of SURVEY 8(d) with this many components instead")
    ap.add_argument("--probes", type=int, default=10)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-recall", action="store_true")
    ap.add_argument("--scan-impl", type=int, default=int(os.environ.get("VB_SCAN_IMPL", "2")),
                    help="0 = LDG.128 streaming scan kernel, 1 = cp.async.bulk (TMA) staged scan kernel, "
                         "2 = library default (list-major batched scan for query batches; per-query: bulk for tables larger "
                         "than L2, LDG for L2-resident ones), 3 = list-major wherever it applies, "
                         "4 = tensor-core filter + exact re-score wherever it applies")
    return ap.parse_args()


# ----------------------------------------------------------------------------- synthetic data + index build (setup, untimed)

def make_dataset(args, device):
    """Synthetic rows (seed 3) and queries (seed 4), generated in slabs to bound temporary memory.

    Default law: 1536-d vectors with low intrinsic dimension -- z ~ N(0, I_L), x = Q z + 0.02 eps with Q a random
    dim x L orthonormal frame (L = --latent-dim, 16).  k-means (the reference's algorithm) then produces balanced
    lists (~0.7x..1.3x of rows/lists), ~10 k candidates per query at probes = 10 and a non-trivial recall@10,
    which is the scan the BASELINE config describes.  --components N selects the Gaussian-mixture law of SURVEY
    8(d) instead; on it the reference's k-means++/Lloyd (CPU oracle and GPU alike) collapses into a few giant
    lists (measured 0/1000/6933 rows per list for 1000 components, 44/1000/60602 for 8192), which turns the
    benchmark into an L2-resident scan of hot lists -- see DESIGN.md section 5."""
    import torch
    if args.components <= 0:
        g = torch.Generator(device=device).manual_seed(3)
        frame = torch.linalg.qr(torch.randn((args.dim, args.latent_dim), generator=g, device=device, dtype=torch.float32))[0]
        rows = torch.empty((args.rows, args.dim), device=device, dtype=torch.float32)
        slab = 65536
        for lo in range(0, args.rows, slab):
            hi = min(args.rows, lo + slab)
            z = torch.randn((hi - lo, args.latent_dim), generator=g, device=device)
            rows[lo:hi] = z @ frame.T + 0.02 * torch.randn((hi - lo, args.dim), generator=g, device=device)
        g2 = torch.Generator(device=device).manual_seed(4)
        zq = torch.randn((args.queries, args.latent_dim), generator=g2, device=device)
        queries = zq @ frame.T + 0.02 * torch.randn((args.queries, args.dim), generator=g2, device=device)
        return rows, queries.contiguous()
    g = torch.Generator(device=device).manual_seed(3)
    comp = torch.randn((args.components, args.dim), generator=g, device=device, dtype=torch.float32)
    rows = torch.empty((args.rows, args.dim), device=device, dtype=torch.float32)
    slab = 65536
    for lo in range(0, args.rows, slab):
        hi = min(args.rows, lo + slab)
        which = torch.randint(0, args.components, (hi - lo,), generator=g, device=device)
        rows[lo:hi] = comp[which] + 0.3 * torch.randn((hi - lo, args.dim), generator=g, device=device)
    g2 = torch.Generator(device=device).manual_seed(4)
    which = torch.randint(0, args.components, (args.queries,), generator=g2, device=device)
    queries = comp[which] + 0.3 * torch.randn((args.queries, args.dim), generator=g2, device=device)
    return rows, queries


def torch_assign(rows, centers, slab=32768):
    """setup-only nearest-centre pass (fp32 matmul, TF32 disabled)"""
    import torch
    out = torch.empty(rows.shape[0], dtype=torch.int64, device=rows.device)
    cn = (centers * centers).sum(1)
    for lo in range(0, rows.shape[0], slab):
        x = rows[lo:lo + slab]
        d = cn[None, :] - 2.0 * (x @ centers.T)
        out[lo:lo + slab] = d.argmin(1)
    return out


def build_index_arrays(args, rows, pv=None):
    """k-means on a sample + assign + group by list.  Uses libvecb200's k-means/assign when built,
    otherwise a torch fp32 Lloyd (setup only; the timed path never touches torch math)."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    n = rows.shape[0]
    g = torch.Generator(device=rows.device).manual_seed(42)
    ns = min(n, max(args.lists * 50, 10000))          # src/ivfbuild.c:448-452
    samp = rows[torch.randperm(n, generator=g, device=rows.device)[:ns]]
    centers = samp[torch.randperm(ns, generator=g, device=rows.device)[:args.lists]].clone()
    how = "torch-lloyd(setup)"
    done = False
    if pv is not None and os.environ.get("VB_BENCH_TORCH_BUILD") != "1":
        try:
            torch.cuda.synchronize()   # torch-made tensors must be complete before the library's stream reads them
            t = pv.Table(pv.VECTOR, args.dim).append(samp)
            pv.synchronize()
            init = pv.kmeans_pp_init(t, pv.L2, args.lists, seed=42)       # InitCenters (src/ivfkmeans.c:23-91)
            c_host, iters = pv.kmeans(t, pv.L2, init, max_iter=500)
            centers = torch.from_numpy(c_host).to(rows.device)
            t.free()
            tr = pv.Table(pv.VECTOR, args.dim).append(rows)
            assign = pv.assign(tr, pv.L2_SQUARED, centers).to(torch.int64)
            tr.free()
            how = f"vb_kmeans({iters} it)+vb_assign"
            done = True
        except pv.VecB200Error as e:
            if e.code != -5:
                raise
    if not done:
        # k-means++ seeding (same algorithm as src/ivfkmeans.c:23-91) then Lloyd, in torch: setup of the CPU arm
        w = torch.full((ns,), float("inf"), device=rows.device)
        cur = int(torch.randint(0, ns, (1,), generator=g, device=rows.device).item())
        sn = (samp * samp).sum(1)
        for i in range(args.lists):
            centers[i] = samp[cur]
            d2 = (sn - 2.0 * (samp @ samp[cur]) + sn[cur]).clamp_(min=0)
            w = torch.minimum(w, d2)
            cur = int(torch.multinomial(w.clamp(min=0) + 1e-30, 1, generator=g).item())
        how = "torch k-means++ + lloyd (setup)"
        for _ in range(10):
            a = torch_assign(samp, centers)
            sums = torch.zeros_like(centers).index_add_(0, a, samp)
            cnt = torch.bincount(a, minlength=args.lists).clamp(min=1).to(torch.float32)
            centers = sums / cnt[:, None]
        assign = torch_assign(rows, centers)
    order = torch.argsort(assign, stable=True)
    counts = torch.bincount(assign, minlength=args.lists)
    offsets = torch.zeros(args.lists + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(counts.cpu(), 0)
    grouped = torch.empty_like(rows)
    slab = 65536
    for lo in range(0, n, slab):
        grouped[lo:lo + slab] = rows[order[lo:lo + slab]]
    lens = counts.cpu().numpy()
    how += f"; list sizes min/mean/max = {int(lens.min())}/{float(lens.mean()):.0f}/{int(lens.max())}"
    return centers.contiguous(), offsets.numpy(), grouped, order.contiguous(), how


# ----------------------------------------------------------------------------- clocks

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append(parts)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def ncu_traffic(args, world, path=None):
    """dram__bytes_read + dram__bytes_write of the list-scan kernel from the committed `ncu --set full` capture
    of this same command (profiles/*_traffic.json); only valid for the shape it was captured on."""
    name = {"ldg": "listscan_traffic.json", "bulk": "listscan_traffic.json", "tile": "listtile_traffic.json", "tc": "listtc_traffic.json"}
    p = os.path.join(ROOT, "profiles", name.get(path, "listscan_traffic.json"))
    default_shape = (args.rows, args.dim, args.lists, args.probes, args.batch, args.components, args.latent_dim) == \
                    (1_000_000, 1536, 1000, 10, 2048, 0, 16)
    if world != 1 or path == "ldg" or not default_shape or not os.path.exists(p):
        return None
    return json.load(open(p))["traffic_bytes"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- CPU arm

def cpu_arm(args, centers, offsets, grouped, ids, queries, seconds):
    """oracle port of GetScanLists + GetScanItems + sort on all host cores; bounded sample"""
    import oracle as O
    cores = os.cpu_count() or 1
    oix = O.Ivf(O.VECTOR, O.L2_SQUARED, centers, offsets, grouped, ids)
    # calibrate on a few queries, then size the sample for ~`seconds`
    t0 = time.perf_counter()
    oix.search_batch(queries[:cores], args.probes, args.k, threads=cores)
    dt = max(time.perf_counter() - t0, 1e-3)
    nq = int(min(len(queries), max(cores, cores * seconds / dt)))
    t0 = time.perf_counter()
    ids_o, dist_o = oix.search_batch(queries[:nq], args.probes, args.k, threads=cores)
    dt = time.perf_counter() - t0
    # single backend figure (amcanparallel = false): one thread
    n1 = max(4, min(nq, int(2.0 / (dt / nq * cores)) if dt > 0 else 4))
    t1 = time.perf_counter()
    oix.search_batch(queries[:n1], args.probes, args.k, threads=1)
    dt1 = time.perf_counter() - t1
    return {"value": nq / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"{nq} queries of the same workload, one query per thread on {cores} threads "
                      f"(oracle port of src/ivfscan.c:47-187 with the reference's compiler flags; no PostgreSQL "
                      f"buffer-manager/fmgr/tuplesort overhead => optimistic)",
            "single_thread_qps": n1 / dt1}, ids_o, dist_o, nq


# ----------------------------------------------------------------------------- main

def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch

    if args.impl == "reference":
        if rank != 0:
            return 0
        dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
        rows, queries = make_dataset(args, dev)
        centers, offsets, grouped, order, how = build_index_arrays(args, rows)
        del rows
        cb, _, _, nq = cpu_arm(args, centers.cpu().numpy(), offsets, grouped.cpu().numpy(), order.cpu().numpy(),
                               queries.cpu().numpy(), max(args.cpu_seconds, 2.0) * max(1, args.steps) / 3.0)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "queries/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1000.0 * args.batch / cb["value"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, how), "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        # rank 0 prints exactly one JSON line on stdout.  NCCL writes its version banner to the process's stdout when
        # the communicator is created (whatever NCCL_DEBUG_FILE says): create it, and run the first collective, with
        # file descriptor 1 pointing at stderr, then put stdout back.
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    import pgvector_b200 as pv
    pv.init(local)
    pv.set_option("scan_impl", args.scan_impl)

    # ---- setup (untimed): data, index, device image
    rows, queries = make_dataset(args, dev)
    torch.cuda.synchronize()
    t_build = time.perf_counter()
    centers, offsets, grouped, order, how = build_index_arrays(args, rows, pv)
    torch.cuda.synchronize()
    how += f"; build {time.perf_counter() - t_build:.2f} s (k-means++ on {min(args.rows, max(args.lists * 50, 10000))} samples, k-means, assign of {args.rows} rows, grouping)"
    del rows
    torch.cuda.empty_cache()
    full_offsets = offsets
    if world > 1:
        # list l lives on rank l % world; the others keep an empty list with the same number
        keep = (torch.arange(args.lists) % world) == rank
        lens = np.diff(offsets)
        sel = torch.zeros(grouped.shape[0], dtype=torch.bool)
        for l in range(args.lists):
            if keep[l]:
                sel[offsets[l]:offsets[l + 1]] = True
        sel = sel.to(dev)
        grouped_local = grouped[sel].contiguous()
        order_local = order[sel].contiguous()
        lens_local = np.where(keep.numpy(), lens, 0)
        offsets = np.zeros(args.lists + 1, dtype=np.int64)
        offsets[1:] = np.cumsum(lens_local)
    else:
        grouped_local, order_local = grouped, order

    ix = pv.IvfflatIndex("vector_l2_ops", args.dim, args.lists)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ix.load(centers, offsets, grouped_local, order_local)
    pv.synchronize()
    upload_s = time.perf_counter() - t0

    stream = torch.cuda.ExternalStream(pv.stream_handle(), device=dev)
    B, k = min(args.batch, args.queries), args.k
    nb = max(1, args.queries // B)
    qbatches = [queries[i * B:(i + 1) * B].contiguous() for i in range(nb)]
    ids_dev = torch.empty((B, k), dtype=torch.int64, device=dev)
    dist_dev = torch.empty((B, k), dtype=torch.float32, device=dev)
    if world > 1:
        g_ids = torch.empty((world, B, k), dtype=torch.int64, device=dev)
        g_dist = torch.empty((world, B, k), dtype=torch.float32, device=dev)

    def step_dev(i):
        ix.search_into(qbatches[i % nb], k, args.probes, ids_dev, dist_dev)
        if world > 1:
            # the one exchange of the list-sharded scan: k (distance, id) pairs per rank, then a k-way merge
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(g_dist, dist_dev)
                dist.all_gather_into_tensor(g_ids, ids_dev)
                d = g_dist.permute(1, 0, 2).reshape(B, world * k)
                ii = g_ids.permute(1, 0, 2).reshape(B, world * k)
                top = torch.topk(d, k, dim=1, largest=False, sorted=True)
                ids_dev.copy_(torch.gather(ii, 1, top.indices))
                dist_dev.copy_(top.values)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    # the clock sampler starts before the warm-up (nvidia-smi needs ~0.3 s to print its first line) and is
    # stopped right after the timed steps: every sample is taken with the scan running
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step_dev(i)
    barrier()
    # untimed extra load (same count on every rank: the steps contain collectives) while nvidia-smi spins up
    # (a step is ~1.6 ms with the batched kernels: a few hundred steps give the 20 ms sampler ~30 lines before the
    # timed region starts; it keeps running through the timed device steps and the end-to-end steps)
    for i in range(300 if args.scan_impl >= 2 else 30):
        step_dev(i)
    barrier()
    pv.prof_enable(True)
    pv.prof_read(pv.PROF_SCAN_ITEMS)
    pv.prof_read(pv.PROF_SCAN_LISTS)
    pv.prof_read(pv.PROF_TOPK)
    l0 = pv.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cand_total = 0
    e0.record(stream)
    for i in range(args.steps):
        step_dev(args.warmup + i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = pv.launch_count() - l0
    scan_ms, scan_n = pv.prof_read(pv.PROF_SCAN_ITEMS)
    lists_ms, lists_n = pv.prof_read(pv.PROF_SCAN_LISTS)
    topk_ms, topk_n = pv.prof_read(pv.PROF_TOPK)
    pv.prof_enable(False)
    # candidates of the last step (same batch size every step; lists differ slightly per batch)
    cand_last = ix.last_candidates()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        c = torch.tensor([cand_last], device=dev, dtype=torch.int64)
        dist.all_reduce(c)
        cand_all = int(c.item())
    else:
        cand_all = cand_last
    qps = args.steps * B / (ms / 1000.0)

    # ---- end to end through the host-buffer C ABI call
    q_host = [torch.empty((B, args.dim), dtype=torch.float32).pin_memory().copy_(qb.cpu()).numpy() for qb in qbatches[:4]]
    ids_h = torch.empty((B, k), dtype=torch.int64).pin_memory().numpy()
    dist_h = torch.empty((B, k), dtype=torch.float64).pin_memory().numpy()

    # Pipelined host path: the H2D copy of step i + 1 (a second stream) overlaps the device work of step i; every
    # step still contains one query upload and one result download, both inside the timed region.
    pipelined = args.dim % 4 == 0 and os.environ.get("VB_BENCH_NO_PIPELINE") != "1"
    e2e_n = [0]
    if pipelined:
        ix.prefetch_queries(q_host[0], 0)

    def step_host(_):
        i = e2e_n[0]
        e2e_n[0] += 1
        if pipelined:
            ix.prefetch_queries(q_host[(i + 1) % len(q_host)], (i + 1) % 2)
            ix.search_prefetched_into(i % 2, k, args.probes, ids_h, dist_h)
        else:
            ix.search_host_into(q_host[i % len(q_host)], k, args.probes, ids_h, dist_h)
        if world > 1:
            with torch.cuda.stream(stream):
                dd = torch.from_numpy(dist_h).to(dev, non_blocking=True).float()
                iid = torch.from_numpy(ids_h).to(dev, non_blocking=True)
                dist.all_gather_into_tensor(g_dist, dd)
                dist.all_gather_into_tensor(g_ids, iid)
                d = g_dist.permute(1, 0, 2).reshape(B, world * k)
                ii = g_ids.permute(1, 0, 2).reshape(B, world * k)
                top = torch.topk(d, k, dim=1, largest=False, sorted=True)
                res = torch.gather(ii, 1, top.indices).cpu()
            stream.synchronize()
            return res

    for i in range(args.warmup):
        step_host(i)
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record(stream)
    for i in range(args.steps):
        step_host(i)
    h1.record(stream)
    barrier()
    ms_h = h0.elapsed_time(h1)
    clocks = sampler.stop() if rank == 0 else None
    # the last end-to-end step against the plain host call on the same batch (local results of this rank)
    last = (e2e_n[0] - 1) % len(q_host)
    chk_ids = np.empty_like(ids_h)
    chk_dist = np.empty_like(dist_h)
    if world == 1:
        got_ids, got_dist = ids_h.copy(), dist_h.copy()
        ix.search_host_into(q_host[last], k, args.probes, chk_ids, chk_dist)
        e2e_matches = bool(np.array_equal(got_ids, chk_ids) and np.array_equal(got_dist, chk_dist))
    else:
        e2e_matches = None
    if world > 1:
        t = torch.tensor([ms_h], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_h = float(t.item())
    e2e_qps = args.steps * B / (ms_h / 1000.0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (list scan), from live CUDA events
    elem_bytes = 4
    cand_per_step = cand_last                         # this rank's candidates in one step
    scan_bytes_per_launch = cand_per_step * args.dim * elem_bytes
    peak, peak_src = measured_peaks()
    scan_avg_ms = scan_ms / max(scan_n, 1)
    achieved = scan_bytes_per_launch / (scan_avg_ms / 1000.0) / 1e9 if scan_avg_ms > 0 else 0.0
    # which kernel the library's choice comes down to for this workload (k <= 40, batched): see vb_set_option
    path = {0: "ldg", 1: "bulk", 3: "tile"}.get(args.scan_impl, "tc" if args.k <= 40 else "tile")
    kernel_name = {"ldg": "scan_kernel", "bulk": "scan_bulk_kernel", "tile": "list_tile_kernel", "tc": "list_tc_kernel"}[path]
    traffic = ncu_traffic(args, world, path)
    roofline = {"bound": "hbm", "kernel": kernel_name + "<vector,L2^2> (GetScanItems list scan)",
                "achieved": achieved,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "bytes_per_launch": scan_bytes_per_launch, "avg_launch_ms": scan_avg_ms,
                "share_of_step": scan_ms / ms if ms > 0 else None,
                "other_kernels_ms_per_step": {"centre_scan": lists_ms / max(lists_n, 1), "topk_select": topk_ms / max(topk_n, 1)},
                "whole_step_algorithmic_gbs": (B * args.lists + cand_all) * args.dim * elem_bytes / (ms / args.steps / 1000.0) / 1e9}
    if path in ("tile", "tc"):
        # list-major kernels read every probed list from HBM ONCE per batch and reuse it for all queries that probe it.
        # 'achieved' keeps SURVEY 8(d)'s definition (one row read per distance, not amortised over the batch), so it
        # exceeds the HBM peak by the reuse factor; the physical roofline is 'dram': ncu DRAM bytes of the launch / its time.
        table_bytes = args.rows * args.dim * elem_bytes // world
        roofline["note"] = ("rows are reused across the queries of a batch: 'achieved' counts one row read per distance (SURVEY 8d, "
                            "not amortised), the DRAM traffic of a launch is one pass over the probed lists (<= %.1f GB)" % (table_bytes / 1e9))
        roofline["table_bytes_per_launch_upper_bound"] = table_bytes
        if traffic and scan_avg_ms > 0:
            dram = traffic / (scan_avg_ms / 1000.0) / 1e9
            roofline["dram"] = {"achieved": dram, "peak": peak, "unit": "GB/s", "frac": dram / peak, "reuse_factor": scan_bytes_per_launch / traffic}
        if path == "tile":
            roofline["fp32_terms_per_s"] = cand_per_step * args.dim / (scan_avg_ms / 1000.0) if scan_avg_ms > 0 else 0.0
        else:
            roofline["certificate_fallback_queries"] = ix.tc_fallbacks()
            roofline["level1_fallback_queries"] = ix.tc_level1_fallbacks()
            # filter level 1 (hi plane of the rows) issues 2 bf16 products per fp32 term, level 2 (both planes) 3
            level = 1 if roofline["level1_fallback_queries"] == 0 else 2
            roofline["filter_level"] = level
            roofline["bf16_mma_tflops_issued"] = (level + 1) * 2.0 * cand_per_step * args.dim / (scan_avg_ms / 1000.0) / 1e12 if scan_avg_ms > 0 else 0.0
            if level == 2:
                roofline.pop("dram", None)     # the committed DRAM-traffic capture is of the level-1 kernel
                roofline["traffic"] = None

    # ---- recall@10 vs exact brute force (GPU exact scan) and CPU baseline
    recall = None
    cpu = None
    if not args.no_recall and world == 1:
        t = pv.Table(pv.VECTOR, args.dim).append(grouped)
        nq_r = min(256, args.queries)
        ex_ids, _ = t.exact_topk(pv.L2_SQUARED, queries[:nq_r].contiguous(), k)
        ex_heap = order[ex_ids.clamp(min=0)]
        got, _ = ix.search(queries[:nq_r].contiguous(), k=k, probes=args.probes)
        hit = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(got.cpu(), ex_heap.cpu()))
        recall = hit / (nq_r * k)
        t.free()
    if not args.no_cpu and world == 1:
        cpu, ids_o, dist_o, nq_c = cpu_arm(args, centers.cpu().numpy(), full_offsets, grouped.cpu().numpy(),
                                            order.cpu().numpy(), queries.cpu().numpy(), args.cpu_seconds)
        # parity of the timed configuration against the oracle on the CPU sample
        got, gd = ix.search(queries[:nq_c].contiguous(), k=k, probes=args.probes)
        cpu["gpu_vs_oracle_id_agreement"] = float((got.cpu().numpy() == ids_o).mean())
        cpu["gpu_vs_oracle_max_rel_dist_err"] = float(np.max(np.abs(gd.cpu().numpy() - dist_o) / np.maximum(np.abs(dist_o), 1e-30)))

    line = {"metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, how), index_upload_s=upload_s,
                           l2_policy=("inputs larger than L2: every step streams ~%d MB of list rows" % (cand_all * args.dim * 4 // 2**20)) if args.scan_impl < 2 else
                                     ("inputs larger than L2: every step reads the probed lists of a %d MB table once" % (args.rows * args.dim * 4 // world // 2**20))),
            "recall_at_10": recall, "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": B * args.dim * 4,
                    "d2h_bytes_per_step": B * k * 16, "ms_per_step": ms_h / args.steps,
                    "call": ("vb_ivf_prefetch_queries (next batch) + vb_ivf_search_prefetched" if pipelined else "vb_ivf_search"),
                    "last_step_equals_plain_call": e2e_matches},
            "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def workload_config(args, how):
    return {"workload": f"IVFFlat L2 {args.rows}x{args.dim} fp32, lists={args.lists}, probes={args.probes}, k={args.k} "
                        f"(BASELINE.json configs[1])",
            "data_law": (f"mixture of {args.components} Gaussians (centres N(0,1), sigma 0.3), seeds 3/4" if args.components > 0 else
                         f"x = Q z + 0.02 eps, z ~ N(0, I_{args.latent_dim}), Q random {args.dim}x{args.latent_dim} orthonormal frame, seeds 3/4"),
            "queries": args.queries, "batch": args.batch,
            "index_build": how, "scan_kernel": {0: "LDG.128 streaming (all scans)", 1: "cp.async.bulk+mbarrier staged (all scans)",
                            3: "list scan: list-major 256x32 fp32x2 register tiles (rows read once per batch); centre scan: 128x128 fp32 tiles",
                            }.get(args.scan_impl, "list scan and probe selection: tcgen05 split-bf16 filter over packed row planes (each probed list read once per "
                                                  "batch; level 1 = hi plane only, level 2 = both planes on certificate failure) + exact fp32 re-score of the "
                                                  "candidates under the certificate threshold; exact list-major kernel as the last resort"),
            "parallelism": "lists sharded l % N, one NCCL all-gather of k results per rank"}


if __name__ == "__main__":
    sys.exit(main())
