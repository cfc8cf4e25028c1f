"""The scan of ONE query (what a backend issues: GetScanLists + GetScanItems for one ORDER BY value, src/ivfscan.c:47-187,
360-414) through the two fused distance + select kernels of csrc/vb_ivf_one.cu: equal to the general (batched) path bit for
bit -- same per-row arithmetic, same tie rules -- and to the oracle within the scan's tolerance."""
import numpy as np
import pytest

import oracle as O
from tests.util import assert_same_neighbours, build_ivf_arrays, f32_to_half_bits, mixture

pytestmark = pytest.mark.gpu
RTOL = 1e-5


@pytest.fixture(scope="module")
def pv():
    import pgvector_b200 as pv
    pv.init(0)
    pv.set_option("scan_impl", 2)
    pv.set_option("one_query", 1)
    O.ivf_set_tie_mode(True)          # (distance, list number) for equal centre distances, like the GPU
    yield pv
    pv.set_option("one_query", 1)
    O.ivf_set_tie_mode(False)


def make_index(pv, opclass, rows, centers, dim=None, assign=None):
    elem, metric, _, _ = pv.OPCLASSES[opclass]
    lists = centers.shape[0]
    if assign is None:
        assign = O.ivf_assign(elem, metric, rows, centers, threads=8, dim=dim)
    grouped, ids, offsets = build_ivf_arrays(rows, assign, lists)
    d = dim if dim is not None else rows.shape[1]
    gix = pv.IvfflatIndex(opclass, d, lists).load(centers, offsets, grouped, ids)
    oix = O.Ivf(elem, metric, centers, offsets, grouped, ids, dim=d)
    return gix, oix


def dataset(pv, opclass, dim, n=6000, lists=20, nq=16, seed=11):
    elem, metric, normalize, _ = pv.OPCLASSES[opclass]
    x, c = mixture(n, dim, lists, seed=seed)
    q, _ = mixture(nq, dim, lists, seed=seed + 1)
    if elem == O.BIT:
        return O.binary_quantize(O.VECTOR, x), O.binary_quantize(O.VECTOR, c), O.binary_quantize(O.VECTOR, q)
    if elem == O.HALFVEC:
        x, c, q = f32_to_half_bits(x), f32_to_half_bits(c), f32_to_half_bits(q)
    if normalize or metric == O.NEG_IP:
        x, c, q = O.l2_normalize(elem, x), O.l2_normalize(elem, c), O.l2_normalize(elem, q)
    return x, c, q


def both_paths(pv, fn):
    """fn() through the fused kernels and through the general path"""
    pv.set_option("one_query", 1)
    a = fn()
    pv.set_option("one_query", 0)
    try:
        b = fn()
    finally:
        pv.set_option("one_query", 1)
    return a, b


OPCLASSES = [("vector_l2_ops", 96), ("vector_l2_ops", 3), ("vector_ip_ops", 40), ("vector_cosine_ops", 40), ("halfvec_l2_ops", 72),
             ("halfvec_cosine_ops", 768), ("vector_l2_ops", 1536), ("bit_hamming_ops", 52), ("bit_hamming_ops", 1024)]


@pytest.mark.parametrize("opclass,dim", OPCLASSES)
def test_search_of_a_few_queries_equals_the_general_path_and_the_oracle(pv, opclass, dim):
    elem, metric, _, _ = pv.OPCLASSES[opclass]
    rows, centers, queries = dataset(pv, opclass, dim)
    gix, oix = make_index(pv, opclass, rows, centers, dim=dim)
    for nq, probes, k in [(1, 4, 10), (3, 1, 5), (16, 20, 40), (1, 7, 1)]:
        (ids, dist), (gi, gd) = both_paths(pv, lambda: gix.search(queries[:nq], k=k, probes=probes))
        if nq * probes < 256 or elem == O.BIT:
            # the general path scans query by query below 256 (query, list) pairs: the same per-row arithmetic
            assert np.array_equal(ids, gi), (nq, probes, k)
            assert np.array_equal(dist, gd), (nq, probes, k)
        else:
            # (list-major there: another summation order)
            assert np.allclose(dist, gd, rtol=RTOL, atol=1e-6)
            assert_same_neighbours(ids, dist, gi, gd, RTOL, min_positional=0.98)
        wi, wd = oix.search_batch(queries[:nq], probes, k, threads=8)
        if elem == O.BIT:
            assert np.array_equal(dist, wd)
            assert np.array_equal(ids, wi)          # ties by scan position on both sides
        else:
            finite = np.isfinite(wd)
            assert np.array_equal(np.isfinite(dist), finite)
            assert np.allclose(dist[finite], wd[finite], rtol=RTOL, atol=1e-6)
            assert_same_neighbours(ids, dist, wi, wd, RTOL, min_positional=0.98)


@pytest.mark.parametrize("opclass,dim", [("vector_l2_ops", 96), ("halfvec_l2_ops", 72), ("bit_hamming_ops", 52)])
def test_scan_lists_then_scan_items_of_one_query(pv, opclass, dim):
    """the two calls the extension glue makes: vb_ivf_scan_lists, then vb_ivf_scan_items with a cap"""
    elem, metric, _, _ = pv.OPCLASSES[opclass]
    rows, centers, queries = dataset(pv, opclass, dim, n=8000, lists=32)
    gix, oix = make_index(pv, opclass, rows, centers, dim=dim)
    for i in range(6):
        for mp in (1, 5, 32, 40):
            (lists, ld), (gl, gld) = both_paths(pv, lambda: gix.scan_lists(queries[i], mp))
            assert np.array_equal(lists, gl) and np.array_equal(ld, gld)
            wl, wd = oix.scan_lists(queries[i], mp)
            n = len(wl)
            assert np.array_equal(lists[0][:n], wl), (i, mp)
            assert np.allclose(ld[0][:n], wd, rtol=RTOL)
            assert np.all(lists[0][n:] == -1) and np.all(np.isinf(ld[0][n:]))
        wl, _ = oix.scan_lists(queries[i], 5)
        for cap in (1, 17, 2048):
            (ids, dist, n), (gi, gd, gn) = both_paths(pv, lambda: gix.scan_items(queries[i], wl, cap=cap))
            assert n == gn and np.array_equal(ids, gi) and np.array_equal(dist, gd), (i, cap)
        wi, wdist, wn = oix.search(queries[i], 5, 0)
        ids, dist, n = gix.scan_items(queries[i], wl, cap=100)
        assert n == wn
        if elem == O.BIT:
            assert np.array_equal(dist, wdist[:100]) and np.array_equal(ids, wi[:100])
        else:
            assert np.allclose(dist, wdist[:100], rtol=RTOL)
            assert (ids == wi[:100]).mean() > 0.97


def test_ties_by_the_thousand_are_settled_by_scan_position(pv):
    """bit(8) rows: nine possible Hamming distances over thousands of candidates, so the k-th place always falls inside a
    run of equal distances; the selection takes the first of them in scan order (as the general path and the oracle do)"""
    rng = np.random.default_rng(5)
    n, lists = 12000, 4
    rows = rng.integers(0, 256, size=(n, 1), dtype=np.uint8)
    centers = np.array([[0x00], [0x0F], [0xF0], [0xFF]], dtype=np.uint8)
    queries = rng.integers(0, 256, size=(8, 1), dtype=np.uint8)
    gix, oix = make_index(pv, "bit_hamming_ops", rows, centers, dim=8)
    for k in (1, 10, 333, 2048):
        for probes in (1, 3):
            (ids, dist), (gi, gd) = both_paths(pv, lambda: gix.search(queries, k=k, probes=probes))
            assert np.array_equal(ids, gi) and np.array_equal(dist, gd), (k, probes)
            wi, wd = oix.search_batch(queries, probes, k, threads=8)
            assert np.array_equal(dist, wd) and np.array_equal(ids, wi), (k, probes)
    # centres at equal distances: the lower list number first
    lists_got, ld = gix.scan_lists(np.array([[0x3C]], dtype=np.uint8), 4)
    assert list(ld[0]) == [4.0, 4.0, 4.0, 4.0] and list(lists_got[0]) == [0, 1, 2, 3]


def test_fewer_candidates_than_k_and_empty_lists(pv):
    rows, centers = mixture(3000, 24, 10, seed=31)
    queries, _ = mixture(4, 24, 10, seed=32)
    assign = O.ivf_assign(O.VECTOR, O.L2_SQUARED, rows, centers, threads=8)
    assign = np.where(assign == 3, 4, assign)          # list 3 is empty
    keep = np.ones(len(rows), dtype=bool)
    idx7 = np.flatnonzero(assign == 7)
    keep[idx7[5:]] = False                             # list 7 holds five rows
    rows, assign = rows[keep], assign[keep]
    gix, oix = make_index(pv, "vector_l2_ops", rows, centers, assign=assign)
    q7 = centers[7:8] + 0.01
    q3 = centers[3:4] + 0.01
    for q, probes in [(q7, 1), (q3, 1), (q3, 2), (queries, 10)]:
        (ids, dist), (gi, gd) = both_paths(pv, lambda: gix.search(q, k=12, probes=probes))
        assert np.array_equal(ids, gi) and np.array_equal(dist, gd)
        wi, wd = oix.search_batch(q, probes, 12, threads=8)
        assert np.array_equal(ids < 0, wi < 0)
        fin = wi >= 0
        assert np.allclose(dist[fin], wd[fin], rtol=RTOL) and np.array_equal(ids[fin], wi[fin])
        assert np.all(np.isinf(dist[~fin]))
    ids, dist = gix.search(q7, k=12, probes=1)
    assert (ids[0] >= 0).sum() == 5
    ids, dist = gix.search(q3, k=12, probes=1)
    assert np.all(ids == -1)


def test_device_resident_queries_take_the_same_kernels(pv):
    torch = pytest.importorskip("torch")
    rows, centers, queries = dataset(pv, "vector_l2_ops", 64, n=5000, lists=16)
    gix, oix = make_index(pv, "vector_l2_ops", rows, centers)
    ids_h, dist_h = gix.search(queries[:7], k=10, probes=3)
    qd = torch.from_numpy(queries[:7]).cuda()
    ids_d, dist_d = gix.search(qd, k=10, probes=3)
    assert np.array_equal(ids_d.cpu().numpy(), ids_h)
    assert np.allclose(dist_d.cpu().numpy(), dist_h, rtol=1e-6)
    before = pv.launch_count()
    gix.search(queries[:1], k=10, probes=3)
    assert pv.launch_count() - before <= 3          # (upload of a vector query is a copy) two fused kernels
