"""CPU tests of the oracle's index-level restatements: tiny-table orderings of the reference's
regression suite, recall floors of its TAP tests, pairing-heap semantics, Elkan == Lloyd."""
import numpy as np
import pytest

import oracle as O
from tests.util import build_ivf_arrays, f32_to_half_bits, load_golden, mixture, parse_vector, recall_at_k

ELEMS = {"vector": O.VECTOR, "halfvec": O.HALFVEC, "bit": O.BIT}
OPC = {"vector_l2_ops": (O.L2_SQUARED, False), "vector_ip_ops": (O.NEG_IP, False), "vector_cosine_ops": (O.NEG_IP, True),
       "vector_l1_ops": (O.L1, False), "halfvec_l2_ops": (O.L2_SQUARED, False), "halfvec_ip_ops": (O.NEG_IP, False),
       "halfvec_cosine_ops": (O.NEG_IP, True), "halfvec_l1_ops": (O.L1, False), "bit_hamming_ops": (O.HAMMING, False),
       "bit_jaccard_ops": (O.JACCARD, False)}
BLOCKS = load_golden("index_orderings.json")["blocks"]


@pytest.mark.parametrize("b", BLOCKS, ids=[x["source"].split("/")[-1] for x in BLOCKS])
def test_reference_tiny_table_orderings(b):
    elem = ELEMS[b["type"]]
    metric, normalize = OPC[b["index"]["opclass"]]
    texts = [v for grp in b["rows"] for v in grp["values"] if v is not None]
    rows = np.stack([parse_vector(t, elem)[0] for t in texts])
    if normalize:
        keep = np.array([O.norm(elem, r) > 0 for r in rows])
        texts = [t for t, kp in zip(texts, keep) if kp]
        rows = O.l2_normalize(elem, rows[keep])
    qry = b["queries"][0]
    qv = parse_vector(qry["query"], elem)[0]
    if normalize:
        qv = O.l2_normalize(elem, qv)
    if b["index"]["am"] == "ivfflat":
        lists = int(b["index"]["options"].split("=")[1])
        centers = rows[:lists].copy()
        assign = O.ivf_assign(elem, metric, rows, centers, dim=b["dim"])
        grouped, ids, offsets = build_ivf_arrays(rows, assign, lists)
        ix = O.Ivf(elem, metric, centers, offsets, grouped, ids, dim=b["dim"])
        got_ids, _, _ = ix.search(qv, lists, 0)
        got = [texts[i] for i in got_ids]
    else:
        g = O.Hnsw(elem, metric, rows, dim=b["dim"])
        e = g.export()
        got_ids, _, _ = g.search(qv, 40)
        got = [texts[e["elem_row"][i]] for i in got_ids]
    want = qry["expected"]
    assert got[:len(want)] == want, (b["source"], got, want)


def test_pairing_heap_scan_lists_is_a_correct_selection():
    rng = np.random.default_rng(0)
    centers = rng.standard_normal((200, 8)).astype(np.float32)
    rows = rng.standard_normal((400, 8)).astype(np.float32)
    assign = O.ivf_assign(O.VECTOR, O.L2_SQUARED, rows, centers)
    grouped, ids, offsets = build_ivf_arrays(rows, assign, 200)
    ix = O.Ivf(O.VECTOR, O.L2_SQUARED, centers, offsets, grouped, ids)
    for mp in (1, 7, 200, 500):
        q = rng.standard_normal(8).astype(np.float32)
        lists, dist = ix.scan_lists(q, mp)
        d_all = O.distance_batch(O.VECTOR, O.L2_SQUARED, q, centers)
        order = np.argsort(d_all, kind="stable")[:min(mp, 200)]
        assert np.array_equal(np.sort(d_all[order]), dist)
        assert set(lists) == set(order)
        assert np.all(np.diff(dist) >= 0)


def test_ivfflat_recall_floors_of_the_reference_tap_tests():
    """test/t/003_ivfflat_vector_build_recall.pl:104-116 on its own data law (uniform 3-d, lists=100, LIMIT 20),
    scaled to 20k rows: probes=1 >= 0.71, 10 >= 0.95, 100 -> 1.0"""
    rng = np.random.default_rng(3)
    rows = rng.random((20000, 3)).astype(np.float32)
    queries = rng.random((20, 3)).astype(np.float32)
    init = O.kmeans_pp_init(O.VECTOR, O.L2, rows[:5000], 100, seed=1)
    centers, _, it = O.kmeans(O.VECTOR, O.L2, rows[:5000], init)
    assert 1 <= it <= 500
    assign = O.ivf_assign(O.VECTOR, O.L2_SQUARED, rows, centers, threads=8)
    grouped, ids, offsets = build_ivf_arrays(rows, assign, 100)
    ix = O.Ivf(O.VECTOR, O.L2_SQUARED, centers, offsets, grouped, ids)
    truth = [O.exact_topk(O.VECTOR, O.L2_SQUARED, q, rows, 20)[0] for q in queries]
    for probes, floor in ((1, 0.71), (10, 0.95), (100, 1.0)):
        got, _ = ix.search_batch(queries, probes, 20, threads=4)
        assert recall_at_k(got, truth) >= floor, probes


def test_hnsw_recall_floor_of_the_reference_tap_tests():
    """test/t/012_hnsw_vector_build_recall.pl:94: 10k x 3-d uniform, m=16, ef_construction=64, ef_search=40 -> >= 0.99"""
    rng = np.random.default_rng(12)
    rows = rng.random((10000, 3)).astype(np.float32)
    queries = rng.random((20, 3)).astype(np.float32)
    g = O.Hnsw(O.VECTOR, O.L2_SQUARED, rows)
    e = g.export()
    truth = [O.exact_topk(O.VECTOR, O.L2_SQUARED, q, rows, 20)[0] for q in queries]
    for ties in (O.TIES_PG, O.TIES_TOTAL):
        ids, _, nd = g.search_batch(queries, 40, 20, ties=ties, threads=4)
        heap = e["elem_row"][ids]
        assert recall_at_k(heap, truth) >= 0.99
        assert np.all(nd > 0)


def test_hnsw_export_import_round_trip():
    rows, _ = mixture(3000, 16, 10, seed=5)
    g = O.Hnsw(O.VECTOR, O.L2_SQUARED, rows)
    e = g.export()
    g2 = O.Hnsw.from_export(O.VECTOR, O.L2_SQUARED, rows[e["elem_row"]], e)
    q, _ = mixture(30, 16, 10, seed=6)
    for ties in (O.TIES_PG, O.TIES_TOTAL):
        a = g.search_batch(q, 50, 10, ties=ties)
        b = g2.search_batch(q, 50, 10, ties=ties)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    # structural invariants of the reference's build: <= 2m / m neighbours, levels consistent
    assert e["nbr0"].shape[1] == 32 and np.all((e["nbr0"] >= -1) & (e["nbr0"] < g.n))
    assert e["levels"][e["entry"]] == e["levels"].max()


def test_elkan_and_lloyd_agree_from_shared_initial_centres():
    """Elkan only prunes distance evaluations (src/ivfkmeans.c:239-245): same assignments as plain Lloyd"""
    for elem, km, conv in ((O.VECTOR, O.L2, lambda x: x), (O.HALFVEC, O.L2, f32_to_half_bits)):
        x, _ = mixture(3000, 12, 15, seed=8)
        x = conv(x)
        init = O.kmeans_pp_init(elem, km, x, 15, seed=4)
        ce, ae, ie = O.kmeans(elem, km, x, init, algo="elkan")
        cl, al, il = O.kmeans(elem, km, x, init, algo="lloyd")
        assert (ae == al).mean() > 0.999
        assert ie == il
    # spherical: unit vectors, centres renormalised
    x, _ = mixture(2000, 10, 8, seed=9)
    x = O.l2_normalize(O.VECTOR, x)
    init = O.kmeans_pp_init(O.VECTOR, O.SPHERICAL, x, 8, seed=2)
    ce, ae, _ = O.kmeans(O.VECTOR, O.SPHERICAL, x, init, algo="elkan")
    cl, al, _ = O.kmeans(O.VECTOR, O.SPHERICAL, x, init, algo="lloyd")
    assert (ae == al).mean() > 0.995
    assert np.allclose(np.linalg.norm(ce, axis=1), 1.0, atol=1e-6)


def test_bit_kmeans_majority_centres():
    """BitUpdateCenter: per-bit majority (> 0.5) of the members (src/ivfutils.c:325-339)"""
    x, _ = mixture(1500, 64, 6, seed=10)
    bits = O.binary_quantize(O.VECTOR, x)
    init = bits[:6].copy()
    c, a, it = O.kmeans(O.BIT, O.HAMMING, bits, init, dim=64)
    un = np.unpackbits(bits, axis=1)
    for j in range(6):
        mem = un[a == j]
        if len(mem):
            want = np.packbits((mem.mean(0) > 0.5).astype(np.uint8))
            assert np.array_equal(c[j], want)


def test_hnsw_iterative_scan_restatement():
    """pgv_hnsw_iter_scan (src/hnswscan.c:62-87, 228-340): batch 0 is the plain scan; resumed batches never repeat an
    element and enumerate everything reachable; past hnsw.max_scan_tuples about that many elements have come
    back (test/t/043_hnsw_iterative_scan.pl: the number of matches of a 1-in-10000 filter is max_scan_tuples / 10000
    +- 2) and the drain is nearest first; the pairing-heap and total-order tie modes agree on float data."""
    rows, _ = mixture(6000, 16, 20, seed=5)
    queries, _ = mixture(8, 16, 20, seed=6)
    g = O.Hnsw(O.VECTOR, O.L2_SQUARED, rows, m=8, ef_construction=32, seed=3)
    for q in queries:
        i0, d0, nd0 = g.search(q, 40, O.TIES_PG)
        ids, dist, batch, nd = g.iter_scan(q, 40, max_scan_tuples=10 ** 9, ties=O.TIES_PG)
        assert np.array_equal(ids[:len(i0)], i0) and np.array_equal(dist[:len(i0)], d0)
        assert len(rows) - 5 <= len(ids) == len(set(ids.tolist())) == nd     # (pruning can leave an element without in-edges)
        for b in range(int(batch.max()) + 1):
            assert np.all(np.diff(dist[batch == b]) >= 0) and (batch == b).sum() <= 40
        ids_t, _, batch_t, nd_t = g.iter_scan(q, 40, max_scan_tuples=10 ** 9, ties=O.TIES_TOTAL)
        if len(np.unique(dist)) == len(dist):       # (two equal fp32 distances among 6000 do happen; the modes may then differ)
            assert np.array_equal(ids, ids_t) and np.array_equal(batch, batch_t) and nd == nd_t
        else:
            assert sorted(ids.tolist()) == sorted(ids_t.tolist())
        for limit in (500, 2000):
            ids2, dist2, batch2, nd2 = g.iter_scan(q, 40, max_scan_tuples=limit, ties=O.TIES_PG)
            assert limit <= len(ids2) == nd2 <= limit + 1500
            assert np.all(np.diff(dist2[batch2 == -1]) >= 0)
            n_same = int((batch2 >= 0).sum())
            assert np.array_equal(ids2[:n_same], ids[:n_same])
