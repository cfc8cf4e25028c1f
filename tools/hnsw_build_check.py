#!/usr/bin/env python3
"""Build an HNSW index on the GPU at a BASELINE shape, time it, and report recall@10 / queries/s of the result.

  python tools/hnsw_build_check.py --config C --rows 1000000
  python tools/hnsw_build_check.py --config E --rows 10000000

Config C: halfvec cosine, 768-d, Gaussian mixture, rows L2-normalised (halfvec_cosine_ops indexes normalised rows,
src/hnswbuild.c:~480 HnswFormIndexValue), ef_search = 100.  Config E: bit(1024) Hamming = binary_quantize of a
1024-d mixture, ef_search = 200.  Prints one JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_rows(cfg, n, nq, dev):
    import torch
    import pgvector_b200 as pv
    dim = 768 if cfg == "C" else 1024
    comps = 1000
    g = torch.Generator(device=dev).manual_seed(3 if cfg == "C" else 6)
    centres = torch.randn((comps, dim), generator=g, device=dev)

    def draw(count, gen):
        out = []
        for lo in range(0, count, 1 << 18):
            m = min(1 << 18, count - lo)
            which = torch.randint(0, comps, (m,), generator=gen, device=dev)
            x = centres[which] + (0.3 if cfg == "C" else 1.0) * torch.randn((m, dim), generator=gen, device=dev)
            if cfg == "C":
                x = torch.nn.functional.normalize(x, dim=1).to(torch.float16)
                x = torch.nn.functional.normalize(x.float(), dim=1).to(torch.float16)
                out.append(x.view(torch.int16))
            else:
                bits = (x > 0).to(torch.uint8).reshape(m, dim // 8, 8)
                w = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], device=dev, dtype=torch.uint8)
                out.append((bits * w).sum(dim=2).to(torch.uint8))
        return torch.cat(out)

    rows = draw(n, g)
    g2 = torch.Generator(device=dev).manual_seed(4 if cfg == "C" else 7)
    queries = draw(nq, g2)
    opclass = "halfvec_cosine_ops" if cfg == "C" else "bit_hamming_ops"
    return rows, queries, opclass, dim


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C", choices=["C", "E"])
    ap.add_argument("--rows", type=int, default=200_000)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--m", type=int, default=16)
    ap.add_argument("--efc", type=int, default=64)
    ap.add_argument("--ef", type=int, default=0)
    args = ap.parse_args()
    import torch
    import pgvector_b200 as pv
    pv.init(0)
    dev = torch.device("cuda", 0)
    rows, queries, opclass, dim = make_rows(args.config, args.rows, args.queries, dev)
    torch.cuda.synchronize()
    ef = args.ef or (100 if args.config == "C" else 200)
    ix = pv.HnswIndex(opclass, dim, m=args.m)
    l0 = pv.launch_count()
    t0 = time.perf_counter()
    ix.build(rows, ef_construction=args.efc, seed=42)
    pv.synchronize()
    build_s = time.perf_counter() - t0
    launches = pv.launch_count() - l0
    k = 10
    ids = torch.empty((args.queries, k), dtype=torch.int64, device=dev)
    dist = torch.empty((args.queries, k), dtype=torch.float32, device=dev)
    nd = torch.empty((args.queries,), dtype=torch.int64, device=dev)
    ix.search_into(queries, k, ef, ids, dist, nd)
    pv.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        ix.search_into(queries, k, ef, ids, dist, nd)
    pv.synchronize()
    qps = reps * args.queries / (time.perf_counter() - t0)
    # recall@10 against the exact scan of the same rows (GPU exact top-k; ids = row numbers = element numbers)
    elem, metric = pv.OPCLASSES[opclass][:2]
    nr = min(500, args.queries)
    t = pv.Table(elem, dim).append(rows)
    ex, exd = t.exact_topk(metric, queries[:nr].contiguous(), k)
    if args.config == "C":
        got, want = ids[:nr].cpu(), ex.cpu()
        hit = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(got, want))
        recall = hit / (nr * k)
    else:   # tie-aware (test/t/020_hnsw_bit_build_recall.pl:85-91)
        recall = float((dist[:nr] <= exd[:, -1:].to(dist.dtype)).float().mean().item())
    g = ix.export()
    deg = float((g["nbr0"] >= 0).sum(axis=1).mean())
    print(json.dumps({"config": args.config, "rows": args.rows, "dim": dim, "opclass": opclass, "m": args.m, "ef_construction": args.efc,
                      "build_s": build_s, "rows_per_s": args.rows / build_s, "build_launches": launches, "ef_search": ef,
                      "recall_at_10": recall, "search_qps": qps, "n_dist_per_query": float(nd.float().mean().item()),
                      "mean_degree_layer0": deg, "duplicates_folded": int((g["dup_of"] >= 0).sum()),
                      "max_level": int(g["levels"].max()), "entry": g["entry"]}))


if __name__ == "__main__":
    main()
