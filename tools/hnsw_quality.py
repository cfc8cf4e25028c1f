#!/usr/bin/env python3
"""Graph quality of the batched GPU build next to the oracle's serial build on the same rows (recall@10 through the
same GPU search), as a function of the batch fraction.  Diagnostic, prints one JSON line per case."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import oracle as O
    import pgvector_b200 as pv
    from tests.util import f32_to_half_bits, mixture, recall_at_k
    pv.init(0)
    cases = [("vector_l2_ops", 32, 8000, 8, 40, 80, 40), ("vector_l2_ops", 48, 20000, 16, 64, 80, 40),
             ("halfvec_cosine_ops", 768, 60000, 16, 64, 100, 300), ("bit_hamming_ops", 1024, 100000, 16, 64, 200, 100)]
    only = sys.argv[1:]
    for ci, (opclass, dim, n, m, efc, ef, comps) in enumerate(cases):
        if only and str(ci) not in only:
            continue
        elem, metric, normalize, _ = pv.OPCLASSES[opclass]
        x, _ = mixture(n, dim, comps, seed=n + dim, sigma=1.0 if elem == O.BIT else 0.3)
        q, _ = mixture(200, dim, comps, seed=n + dim + 1, sigma=1.0 if elem == O.BIT else 0.3)
        if elem == O.BIT:
            x, q = O.binary_quantize(O.VECTOR, x), O.binary_quantize(O.VECTOR, q)
        elif elem == O.HALFVEC:
            x, q = f32_to_half_bits(x), f32_to_half_bits(q)
        if normalize:
            x, q = O.l2_normalize(elem, x), O.l2_normalize(elem, q)
        k = 10
        truth = [O.exact_topk(elem, metric, qq, x, k, dim=dim) for qq in q]

        def score(ids, dist):
            if elem == O.BIT:
                return float(np.mean([np.mean(d <= t[1][-1]) for d, t in zip(dist, truth)]))
            return recall_at_k(ids, [t[0] for t in truth])

        out = {"case": f"{opclass} {n}x{dim} m={m} efc={efc} ef={ef}"}
        t0 = time.perf_counter()
        ob = O.Hnsw(elem, metric, x, m=m, ef_construction=efc, seed=7, dim=dim)
        out["oracle_build_s"] = time.perf_counter() - t0
        g = ob.export()
        gi = pv.HnswIndex(opclass, dim, m=m).load(x[g["elem_row"]], g["levels"], g["nbr0"], g["upper_off"], g["upper"], g["entry"])
        ids, dist, nd = gi.search(q, k=k, ef_search=ef)
        out["serial_build"] = {"recall": score(g["elem_row"][np.maximum(ids, 0)], dist), "n_dist": float(nd.mean()),
                               "degree": float((g["nbr0"] >= 0).sum(axis=1).mean())}
        for frac in (8, 64, 256, 1 << 30):
            pv.set_option("hnsw_build_fraction", frac)
            t0 = time.perf_counter()
            gb = pv.HnswIndex(opclass, dim, m=m).build(x, ef_construction=efc, seed=7, levels=g["levels"] if len(g["levels"]) == n else None)
            dt = time.perf_counter() - t0
            ids, dist, nd = gb.search(q, k=k, ef_search=ef)
            ge = gb.export()
            out[f"gpu_build_fraction_{frac if frac < 1 << 30 else 'sequential'}"] = {
                "recall": score(ids, dist), "n_dist": float(nd.mean()), "build_s": dt, "degree": float((ge["nbr0"] >= 0).sum(axis=1)[ge["dup_of"] < 0].mean())}
            if n > 30000 and frac == 256:
                break
        pv.set_option("hnsw_build_fraction", 64)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
