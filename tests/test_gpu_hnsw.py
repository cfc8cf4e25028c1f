"""GPU parity of the HNSW search path (GetScanItems / HnswSearchLayer, src/hnswscan.c:25-56,
src/hnswutils.c:824-987) against the oracle on graphs built by the oracle's restatement of the
reference's in-memory build."""
import numpy as np
import pytest

import oracle as O
from tests.util import f32_to_half_bits, load_golden, mixture, parse_vector, recall_at_k

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def pv():
    import pgvector_b200 as pv
    pv.init(0)
    return pv


def build_pair(pv, opclass, rows, dim=None, m=16, efc=64, seed=7):
    elem, metric, _, _ = pv.OPCLASSES[opclass]
    og = O.Hnsw(elem, metric, rows, m=m, ef_construction=efc, seed=seed, dim=dim)
    g = og.export()
    erows = rows[g["elem_row"]]
    d = dim if dim is not None else rows.shape[1]
    gi = pv.HnswIndex(opclass, d, m=m).load(erows, g["levels"], g["nbr0"], g["upper_off"], g["upper"], g["entry"])
    return og, gi, g


@pytest.fixture(scope="module")
def l2_graph(pv):
    rows, _ = mixture(20000, 48, 50, seed=21)
    queries, _ = mixture(300, 48, 50, seed=22)
    og, gi, g = build_pair(pv, "vector_l2_ops", rows)
    return og, gi, g, rows, queries


@pytest.mark.parametrize("ef,k", [(1, 1), (10, 10), (40, 10), (100, 100), (200, 50)])
def test_l2_search_matches_total_order_oracle(l2_graph, ef, k):
    og, gi, g, rows, queries = l2_graph
    ids, dist, nd = gi.search(queries, k=k, ef_search=ef)
    wi, wd, wnd = og.search_batch(queries, ef, k, ties=O.TIES_TOTAL, threads=8)
    finite = wi >= 0
    assert np.array_equal(ids >= 0, finite)
    assert np.allclose(dist[finite], wd[finite], rtol=RTOL)
    # fp32 summation order can flip a near tie and send the walk elsewhere; it must be rare
    same_q = np.all(ids == wi, axis=1)
    assert same_q.mean() > 0.97, same_q.mean()
    assert np.array_equal(nd[same_q], wnd[same_q])        # identical walks evaluate identical distance counts
    assert np.all(np.diff(dist, axis=1)[finite[:, 1:]] >= 0)


def test_recall_equals_reference_tie_mode(l2_graph):
    """total-order mode (GPU) vs PostgreSQL pairing-heap mode (reference semantics): no ties in
    float data => same results; recall identical."""
    og, gi, g, rows, queries = l2_graph
    ids, _, _ = gi.search(queries, k=10, ef_search=40)
    pg_ids, _, _ = og.search_batch(queries, 40, 10, ties=O.TIES_PG, threads=8)
    truth = [O.exact_topk(O.VECTOR, O.L2_SQUARED, q, rows, 10)[0] for q in queries]
    r_gpu, r_pg = recall_at_k(ids, truth), recall_at_k(pg_ids, truth)
    assert abs(r_gpu - r_pg) < 2e-3
    assert r_gpu > 0.5       # clustered 48-d data at ef=40; the reference's own floor (>= 0.99 on uniform 3-d, 012:94) is in test_uniform_3d_recall_floor


def test_uniform_3d_recall_floor(pv):
    """test/t/012_hnsw_vector_build_recall.pl:94: 10k x 3-d uniform, defaults, ef_search = 40 -> recall >= 0.99"""
    rng = np.random.default_rng(12)
    rows = rng.random((10000, 3)).astype(np.float32)
    queries = rng.random((50, 3)).astype(np.float32)
    og, gi, g = build_pair(pv, "vector_l2_ops", rows)
    ids, _, _ = gi.search(queries, k=20, ef_search=40)
    heap = g["elem_row"][ids]
    truth = [O.exact_topk(O.VECTOR, O.L2_SQUARED, q, rows, 20)[0] for q in queries]
    assert recall_at_k(heap, truth) >= 0.99


@pytest.mark.parametrize("opclass,dim,n", [("vector_ip_ops", 32, 6000), ("vector_cosine_ops", 32, 6000), ("vector_l1_ops", 16, 5000),
                                           ("halfvec_l2_ops", 40, 6000), ("halfvec_cosine_ops", 768, 3000)])
def test_float_opclasses(pv, opclass, dim, n):
    elem, metric, normalize, _ = pv.OPCLASSES[opclass]
    x, _ = mixture(n, dim, 30, seed=31)
    q, _ = mixture(100, dim, 30, seed=32)
    if elem == O.HALFVEC:
        x, q = f32_to_half_bits(x), f32_to_half_bits(q)
    if normalize:
        x, q = O.l2_normalize(elem, x), O.l2_normalize(elem, q)   # HnswNormValue (src/hnswscan.c:109-110)
    og, gi, g = build_pair(pv, opclass, x)
    ids, dist, nd = gi.search(q, k=10, ef_search=60)
    wi, wd, wnd = og.search_batch(q, 60, 10, ties=O.TIES_TOTAL, threads=8)
    assert np.allclose(dist, wd, rtol=RTOL, atol=1e-6)
    assert np.all(ids == wi, axis=1).mean() > 0.95


@pytest.mark.parametrize("opclass,dim", [("bit_hamming_ops", 52), ("bit_hamming_ops", 1024), ("bit_jaccard_ops", 256)])
def test_bit_opclasses_are_bit_exact(pv, opclass, dim):
    """integer metrics: distances AND ids identical to the total-order oracle, every query"""
    elem, metric, _, _ = pv.OPCLASSES[opclass]
    x, _ = mixture(8000, dim, 40, seed=41)
    q, _ = mixture(200, dim, 40, seed=42)
    rows, queries = O.binary_quantize(O.VECTOR, x), O.binary_quantize(O.VECTOR, q)
    og, gi, g = build_pair(pv, opclass, rows, dim=dim)
    # duplicates share an element (src/hnswbuild.c:343-364)
    assert g["n_heaptids"].sum() == 8000
    ids, dist, nd = gi.search(queries, k=20, ef_search=100)
    wi, wd, wnd = og.search_batch(queries, 100, 20, ties=O.TIES_TOTAL, threads=8)
    assert np.array_equal(dist, wd)
    assert np.array_equal(ids, wi)
    assert np.array_equal(nd, wnd)
    # against the pairing-heap tie order (reference semantics) only the recall is comparable
    pg_ids, pg_d, _ = og.search_batch(queries, 100, 20, ties=O.TIES_PG, threads=8)
    erows = rows[g["elem_row"]]
    truth = [O.exact_topk(elem, metric, qq, erows, 20, dim=dim)[1] for qq in queries]
    # tie-aware recall like test/t/020_hnsw_bit_build_recall.pl:85-91: count results within the true k-th distance
    def tie_recall(d):
        return np.mean([np.mean(di <= t[-1]) for di, t in zip(d, truth)])
    assert abs(tie_recall(dist) - tie_recall(pg_d)) < 0.06


def test_reference_hnsw_orderings(pv):
    """tiny-table orderings of test/expected/hnsw_*.out"""
    blocks = [b for b in load_golden("index_orderings.json")["blocks"] if b["index"]["am"] == "hnsw"]
    assert len(blocks) >= 9
    ELEMS = {"vector": O.VECTOR, "halfvec": O.HALFVEC, "bit": O.BIT}
    for b in blocks:
        elem = ELEMS[b["type"]]
        opclass = b["index"]["opclass"]
        _, metric, normalize, _ = pv.OPCLASSES[opclass]
        texts = [v for grp in b["rows"] for v in grp["values"] if v is not None]
        rows = np.stack([parse_vector(t, elem)[0] for t in texts])
        if normalize:
            keep = np.array([O.norm(elem, r) > 0 for r in rows])
            texts = [t for t, kp in zip(texts, keep) if kp]
            rows = O.l2_normalize(elem, rows[keep])
        og, gi, g = build_pair(pv, opclass, rows, dim=b["dim"])
        qry = b["queries"][0]
        qv = parse_vector(qry["query"], elem)[0]
        if normalize:
            qv = O.l2_normalize(elem, qv)
        ids, dist, _ = gi.search(qv, k=len(texts), ef_search=40)
        got = [texts[g["elem_row"][i]] for i in ids[0] if i >= 0]
        assert got[:len(qry["expected"])] == qry["expected"], (b["source"], got)


def test_empty_and_single_element_index(pv):
    gi = pv.HnswIndex("vector_l2_ops", 3)
    gi.load(np.zeros((0, 3), np.float32), np.zeros(0, np.int32), np.zeros((0, 32), np.int32), np.zeros(0, np.int64),
            np.zeros((0, 16), np.int32), -1)
    ids, dist, nd = gi.search(np.ones((2, 3), np.float32), k=5, ef_search=10)
    assert np.all(ids == -1)
    gi2 = pv.HnswIndex("vector_l2_ops", 3)
    nbr0 = np.full((1, 32), -1, np.int32)
    gi2.load(np.array([[1, 2, 3]], np.float32), np.zeros(1, np.int32), nbr0, np.full(1, -1, np.int64), np.zeros((0, 16), np.int32), 0)
    ids, dist, nd = gi2.search(np.array([[1, 2, 4]], np.float32), k=3, ef_search=10)
    assert list(ids[0]) == [0, -1, -1] and dist[0][0] == 1.0 and nd[0] == 1


def test_visited_set_with_duplicate_neighbours_and_many_queries(pv):
    """the per-warp visited table (bucketed, lanes of one expansion arbitrated in registers): neighbour lists that name the
    same element twice (a second occurrence is "visited", src/hnswutils.c:907-921), every query of a batch much larger than
    the resident warps (queries are handed out dynamically) -- ids, distances and the tuples counter equal the oracle's"""
    x, _ = mixture(6000, 64, 40, seed=51)
    q, _ = mixture(6000, 64, 40, seed=52)
    rows, queries = O.binary_quantize(O.VECTOR, x), O.binary_quantize(O.VECTOR, q)
    og0 = O.Hnsw(O.BIT, O.HAMMING, rows, m=16, ef_construction=64, seed=5, dim=64)
    g = og0.export()
    nbr0 = g["nbr0"].copy()
    rng = np.random.default_rng(53)
    for i in rng.choice(nbr0.shape[0], nbr0.shape[0] // 3, replace=False):
        cnt = int((nbr0[i] >= 0).sum())
        if cnt >= 3:
            a, b = rng.choice(cnt, 2, replace=False)
            nbr0[i, a] = nbr0[i, b]                     # the same neighbour twice in one list
            if cnt >= 8:
                nbr0[i, cnt - 1] = nbr0[i, 0]
    g2 = dict(g, nbr0=nbr0)
    erows = rows[g["elem_row"]]
    og = O.Hnsw.from_export(O.BIT, O.HAMMING, erows, g2, dim=64)
    gi = pv.HnswIndex("bit_hamming_ops", 64, m=16).load(erows, g["levels"], nbr0, g["upper_off"], g["upper"], g["entry"])
    ids, dist, nd = gi.search(queries, k=10, ef_search=64)
    wi, wd, wnd = og.search_batch(queries, 64, 10, ties=O.TIES_TOTAL, threads=8)
    assert np.array_equal(dist, wd)
    assert np.array_equal(ids, wi)
    assert np.array_equal(nd, wnd)


def test_visited_table_grows_when_a_search_fills_it(pv):
    """a random graph (every list names 2m random elements, almost all of them fresh): a search visits ~32 elements per
    expansion, the table sized from ef x m (4096 slots) fills beyond three quarters, the launch is repeated with a larger
    one and the results are those of the oracle on the same graph"""
    rows, _ = mixture(20000, 16, 20, seed=61)
    queries, _ = mixture(64, 16, 20, seed=62)
    n, m = rows.shape[0], 16
    nbr0 = np.random.default_rng(63).integers(0, n, size=(n, 2 * m)).astype(np.int32)
    g = dict(levels=np.zeros(n, np.int32), nbr0=nbr0, upper_off=np.full(n, -1, np.int64), upper=np.zeros((0, m), np.int32),
             entry=0, entry_level=0, m=m)
    og = O.Hnsw.from_export(O.VECTOR, O.L2_SQUARED, rows, g)
    gi = pv.HnswIndex("vector_l2_ops", 16, m=m).load(rows, g["levels"], nbr0, g["upper_off"], g["upper"], 0)
    for ef in (127, 100):      # (the grown size is remembered per ef_search: both calls go through the repeat)
        ids, dist, nd = gi.search(queries, k=50, ef_search=ef)
        wi, wd, wnd = og.search_batch(queries, ef, 50, ties=O.TIES_TOTAL, threads=8)
        assert np.allclose(dist, wd, rtol=RTOL)
        same_q = np.all(ids == wi, axis=1)
        assert same_q.mean() > 0.9, same_q.mean()
        assert np.array_equal(nd[same_q], wnd[same_q])
        assert nd.max() > 3072        # (a search did outgrow 4096 slots, three quarters usable)


# ------------------------------------------------------------------------------------------------ iterative scan

def scan_all(gi, queries, ef, max_scan_tuples, max_batches=10 ** 6):
    """drive an HnswScan to exhaustion: per query the concatenated (ids, distances, batch sizes)"""
    nq = len(queries)
    ids = [[] for _ in range(nq)]
    dist = [[] for _ in range(nq)]
    sizes = [[] for _ in range(nq)]
    with gi.iterative_scan(queries, ef_search=ef, max_scan_tuples=max_scan_tuples) as sc:
        for _ in range(max_batches):
            bi, bd, cnt = sc.next_batch()
            if not cnt.any():
                break
            for q in range(nq):
                c = int(cnt[q])
                if c:
                    ids[q].extend(bi[q, :c].tolist())
                    dist[q].extend(bd[q, :c].tolist())
                    sizes[q].append(c)
                assert np.all(bi[q, c:] == -1)
        tuples = sc.tuples()
    return ids, dist, sizes, tuples


@pytest.mark.parametrize("ef,max_tuples", [(40, 10 ** 9), (40, 3000), (10, 500), (100, 20000)])
def test_iterative_scan_matches_oracle(l2_graph, ef, max_tuples):
    """hnsw.iterative_scan = relaxed_order (src/hnswscan.c:62-87, 228-340): the element sequence of every batch
    (GetScanItems, then ResumeScanItems from the discarded candidates on the same visited set, then the drain past
    hnsw.max_scan_tuples) equals the total-order oracle's, element for element, with the same tuples counter."""
    og, gi, g, rows, queries = l2_graph
    queries = queries[:24]
    ids, dist, sizes, tuples = scan_all(gi, queries, ef, max_tuples)
    same = prefix_same = 0
    overlap = []
    for q in range(len(queries)):
        wi, wd, wb, wt = og.iter_scan(queries[q], ef, max_scan_tuples=max_tuples, ties=O.TIES_TOTAL)
        got = np.array(ids[q])
        assert len(set(ids[q])) == len(ids[q])                      # an element is returned once
        # every batch is sorted by distance, the drain as a whole too
        at = 0
        for c in sizes[q]:
            assert np.all(np.diff(dist[q][at:at + c]) >= 0)
            at += c
        if len(got) == len(wi) and np.array_equal(got, wi):
            same += 1
            assert np.allclose(dist[q], wd, rtol=RTOL)
            assert int(tuples[q]) == wt
            # batch boundaries: the searched batches as the oracle cut them, the drain in ef-sized pieces
            searched = [int((wb == b).sum()) for b in range(int(wb.max()) + 1)] if len(wb) else []
            assert sizes[q][:len(searched)] == searched
        # thousands of elements ordered by fp32 distances: a summation-order flip between two nearly equal ones is
        # expected somewhere in a long scan, so long scans are compared as sets / prefixes as well
        n3 = int(np.sum((wb >= 0) & (wb < 3)))                      # (the searched batches 0..2; -1 marks the drain)
        prefix_same += int(np.array_equal(got[:n3], wi[:n3]))
        overlap.append(len(set(ids[q]) & set(wi.tolist())) / max(1, len(wi)))
        if len(got) == len(wi):
            assert np.allclose(np.sort(dist[q]), np.sort(wd), rtol=1e-4)
        # the first batch is the plain scan
        pi, pd, _ = gi.search(queries[q], k=ef, ef_search=ef)
        n0 = sizes[q][0]
        assert np.array_equal(got[:n0], pi[0][:n0])
        if max_tuples >= 10 ** 9:
            assert len(rows) - 5 <= len(got) == len(wi)             # everything reachable is enumerated
        else:
            # test/t/043_hnsw_iterative_scan.pl: about max_scan_tuples elements come back in total
            assert max_tuples <= len(got) <= max_tuples + 100 * ef + 200
    assert prefix_same >= len(queries) - 2, prefix_same              # the first three batches, element for element
    assert np.mean(overlap) > 0.98, np.mean(overlap)
    if max_tuples <= 500:
        assert same >= len(queries) - 2, same                       # short scans are identical outright


def test_iterative_scan_strict_order_and_limit(l2_graph):
    """strict_order drops elements nearer than one already returned (src/hnswscan.c:316-322); a LIMIT stops pulling"""
    og, gi, g, rows, queries = l2_graph
    with gi.iterative_scan(queries[:1], ef_search=20, max_scan_tuples=2000) as sc:
        got = sc.tuples_of(0, strict=True)
    d = np.array([x[1] for x in got])
    assert len(got) > 20 and np.all(np.diff(d) >= 0)
    wi, wd, _, _ = og.iter_scan(queries[0], 20, max_scan_tuples=2000, ties=O.TIES_TOTAL)
    keep, prev = [], -np.inf
    for e, x in zip(wi, wd):
        if x >= prev:
            keep.append(int(e))
            prev = x
    assert [x[0] for x in got] == keep
    with gi.iterative_scan(queries[:1], ef_search=20, max_scan_tuples=2000) as sc:
        assert [x[0] for x in sc.tuples_of(0, limit=55)] == [int(e) for e in wi[:55]]


def test_iterative_scan_small_and_filtered(pv):
    """fewer elements than ef_search; a filter that keeps 1 row in 50 finds its LIMIT through resumed batches
    (the WHERE i % 10000 = 0 ... LIMIT 11 shape of test/t/043_hnsw_iterative_scan.pl)"""
    rng = np.random.default_rng(5)
    rows = rng.random((30, 3)).astype(np.float32)
    og, gi, g = build_pair(pv, "vector_l2_ops", rows, m=4, efc=16)
    ids, dist, sizes, tuples = scan_all(gi, rows[:2], 40, 20000)
    for q in range(2):
        wi, _, _, wt = og.iter_scan(rows[q], 40, ties=O.TIES_TOTAL)
        assert ids[q] == wi.tolist() and int(tuples[q]) == wt
    rows = rng.random((20000, 3)).astype(np.float32)
    og, gi, g = build_pair(pv, "vector_l2_ops", rows, m=8, efc=32)
    er = g["elem_row"]
    q = rows[7]
    with gi.iterative_scan(q, ef_search=40, max_scan_tuples=20000) as sc:
        hits = []
        while len(hits) < 11:
            bi, bd, cnt = sc.next_batch()
            if cnt[0] == 0:
                break
            hits += [int(er[e]) for e in bi[0, :cnt[0]] if er[e] % 50 == 0]
    assert len(hits) >= 11
    # the nearest filtered rows are found (relaxed order: compare as sets against the exact answer)
    d = ((rows[::50] - q) ** 2).sum(1)
    exact = set((np.argsort(d)[:5] * 50).tolist())
    assert len(exact & set(hits)) >= 4
