"""GPU parity (through the C ABI): sparsevec distance functions, l2_norm / l2_normalize and the exact scan over a
resident CSR table vs the oracle (oracle/pgv_sparse.c) and vs the reference's known-answer outputs
(test/expected/sparsevec.out, hnsw_sparsevec.out)."""
import math

import numpy as np
import pytest

import oracle as O
from tests.test_oracle_sparse import KAT, METRIC, OPS, ORDERINGS, expect_float, random_sparse, sv

pytestmark = pytest.mark.gpu

RTOL = 1e-5   # north_star: L2 / IP / cosine distances within 1e-5 relative


@pytest.fixture(scope="module")
def pv():
    import pgvector_b200 as pv
    pv.init(0)
    return pv


@pytest.mark.parametrize("case", KAT, ids=[c["source"].split("/")[-1] for c in KAT])
def test_known_answers_on_gpu(pv, case):
    S = pv.sparsevec
    fn = case["fn"]
    args = [S.SparseVector.from_text(a) for a in case["args"]]
    if fn in METRIC:
        a, b = args
        if case["error"]:
            with pytest.raises(ValueError) as e:
                S.distance_batch(METRIC[fn], b, a)
            assert str(e.value) == case["error"]
            return
        got = S.distance_batch(METRIC[fn], b, a)[0]      # row = first argument, query = second
        want = expect_float(case["expected"])
        assert (math.isnan(got) and math.isnan(want)) or got == want, (case, got)
    elif fn == "l2_norm":
        got = S.l2_norm(args[0])[0]
        if case["real"]:
            assert np.float32(got) == np.float32(expect_float(case["expected"]))
        else:
            assert got == expect_float(case["expected"])
    elif fn == "l2_normalize":
        got = S.l2_normalize(args[0]).row(0)
        want = S.SparseVector.from_text(case["expected"])
        assert np.array_equal(got.indices, want.indices) and np.array_equal(got.values, want.values), (got, want)
    else:
        pytest.fail(f"unhandled {fn}")


def _close(got, truth):
    got, truth = np.asarray(got), np.asarray(truth)
    nan = np.isnan(truth)
    assert np.array_equal(np.isnan(got), nan)
    assert np.all(np.abs(got - truth)[~nan] <= RTOL * np.maximum(np.abs(truth[~nan]), 1e-30))


def _ip_mag(v, q):
    _, ia, ib = np.intersect1d(v.indices, q.indices, assume_unique=True, return_indices=True)
    return float(np.sum(np.abs(v.values[ia].astype(np.float64) * q.values[ib])))


@pytest.mark.parametrize("dim,row_nnz,q_nnz,n", [(64, 20, 30, 500), (10_000, 120, 300, 4000), (1_000_000_000, 1000, 1000, 600),
                                                 (1_000_000, 16_000, 16_000, 40), (3000, 5, 2, 2000), (100_000, 200, 0, 300)])
def test_distance_batch_matches_oracle(pv, dim, row_nnz, q_nnz, n):
    S = pv.sparsevec
    rng = np.random.default_rng(dim % 97 + n)
    q = random_sparse(rng, dim, q_nnz) if q_nnz else S.SparseVector(dim)
    rows = []
    for r in range(n):
        v = random_sparse(rng, dim, int(rng.integers(0, row_nnz + 1)))
        if q.nnz and r % 2 == 0:        # share some indices with the query, otherwise huge dims never match
            take = rng.choice(q.nnz, size=min(q.nnz, max(1, v.nnz // 2 + 1)), replace=False)
            idx = np.concatenate([v.indices, q.indices[take]])
            val = np.concatenate([v.values, rng.standard_normal(take.size).astype(np.float32)])
            _, first = np.unique(idx, return_index=True)
            first = first[:S.SPARSEVEC_MAX_NNZ]
            v = S.SparseVector(dim, idx[first], val[first])
        rows.append(v)
    rows[0] = S.SparseVector(dim)                        # an empty row
    rows[1] = S.SparseVector(dim, q.indices, q.values)    # the query itself
    R = S.SparseRows.from_vectors(rows, dim)
    for m in (O.L2_SQUARED, O.L2, O.L1, O.COSINE, O.IP, O.NEG_IP):
        got = S.distance_batch(m, q, R)
        truth = np.array([O.sparse_distance(m, sv(v), sv(q), f64=True) for v in rows])
        if m in (O.IP, O.NEG_IP):
            # products cancel: bound by the sum of |a_i b_i| over the matches (what the fp32 sum can lose)
            mag = np.array([_ip_mag(v, q) for v in rows])
            assert np.all(np.abs(got - truth) <= RTOL * np.maximum(mag, 1e-30))
        elif m == O.COSINE:
            # 1 - similarity: the similarity is good to 1e-5 relative, the distance to 1e-5 absolute
            nan = np.isnan(truth)
            assert np.array_equal(np.isnan(got), nan) and np.all(np.abs(got - truth)[~nan] <= RTOL)
        else:
            _close(got, truth)
        # the oracle's fp32 restatement is within the same tolerance of the truth, so the two agree to 2 RTOL
        ref = O.sparse_distance_batch(m, sv(q), R.row_off, R.idx, R.val)
        if m not in (O.IP, O.NEG_IP, O.COSINE):
            ok = np.isnan(ref) | (np.abs(got - ref) <= 2 * RTOL * np.maximum(np.abs(ref), 1e-30))
            assert ok.all()
    # identical vectors are at distance exactly 0 (matched terms only, no cancellation)
    assert S.distance_batch(O.L2, q, R)[1] == 0.0 and S.distance_batch(O.L1, q, R)[1] == 0.0
    # NULL query
    assert np.array_equal(S.distance_batch(O.L2, None, R), np.zeros(n))


def test_norm_and_normalize_match_oracle(pv):
    S = pv.sparsevec
    rng = np.random.default_rng(11)
    dim = 50_000
    rows = [random_sparse(rng, dim, int(rng.integers(0, 400))) for _ in range(300)]
    rows.append(S.SparseVector(dim, [3, 9, 11, 20], [3e37, 3e-37, 4e37, 4e-37]))   # quotients that round to zero are dropped
    rows.append(S.SparseVector(dim))
    R = S.SparseRows.from_vectors(rows, dim)
    norms = S.l2_norm(R)
    want = np.array([O.sparse_l2_norm(sv(v)) for v in rows])
    assert np.all(np.abs(norms - want) <= 1e-12 * np.maximum(want, 1e-300))
    N = S.l2_normalize(R)
    for r, v in enumerate(rows):
        wi, wv = O.sparse_l2_normalize(sv(v))
        g = N.row(r)
        assert np.array_equal(g.indices, wi)
        # quotient by an fp64 norm that may differ in the last bits: one float ulp
        assert np.all(np.abs(g.values - wv) <= 1.2e-7 * np.abs(wv))
    assert N.row(len(rows) - 2).nnz == 2
    # (float_overflow_error() of src/sparsevec.c:1107 is unreachable: |x| <= norm, so no quotient exceeds 1)


@pytest.mark.parametrize("block", ORDERINGS, ids=[b["index"]["opclass"] for b in ORDERINGS])
def test_tiny_orderings_through_the_exact_scan(pv, block):
    S = pv.sparsevec
    vals = [v for grp in block["rows"] for v in grp["values"] if v is not None]
    t = S.SparseTable(block["dim"]).append([S.SparseVector.from_text(v) for v in vals])
    assert t.rows == len(vals)
    for qd in block["queries"]:
        ids, dist = t.exact_topk(OPS[qd["op"]], [S.SparseVector.from_text(qd["query"])], len(vals))
        got = [vals[i] for i, d in zip(ids[0], dist[0]) if not math.isnan(d)]
        assert got == qd["expected"]
    t.free()


@pytest.mark.parametrize("metric", [O.L2, O.NEG_IP, O.COSINE, O.L1])
def test_exact_topk_matches_oracle(pv, metric):
    S = pv.sparsevec
    rng = np.random.default_rng(metric + 20)
    dim, n, nq, k = 30_000, 6000, 48, 10
    rows = [random_sparse(rng, dim, int(rng.integers(1, 150))) for _ in range(n)]
    queries = [random_sparse(rng, dim, int(rng.integers(1, 200))) for _ in range(nq)]
    t = S.SparseTable(dim)
    t.append(rows[:2500]).append(rows[2500:])       # two appends: offsets are rebased on the device
    assert t.rows == n and t.nnz == sum(v.nnz for v in rows)
    ids, dist = t.exact_topk(metric, queries, k)
    R = S.SparseRows.from_vectors(rows, dim)
    for qi, q in enumerate(queries):
        d = O.sparse_distance_batch(metric, sv(q), R.row_off, R.idx, R.val)
        order = np.argsort(d, kind="stable")[:k]
        # ids agree except across candidates closer than the tolerance
        if not np.array_equal(ids[qi], order):
            cut = d[order[-1]]
            for i in set(ids[qi].tolist()) ^ set(order.tolist()):
                assert abs(d[i] - cut) <= 2 * RTOL * max(abs(cut), 1e-30)
        assert np.all(np.abs(dist[qi] - d[ids[qi]]) <= 2 * RTOL * np.maximum(np.abs(d[ids[qi]]), 1e-6))
    t.free()


def test_error_texts(pv):
    S = pv.sparsevec
    with pytest.raises(ValueError) as e:
        S.l2_distance(S.SparseVector(3, [0], [1.0]), [S.SparseVector(2, [0], [1.0])])
    assert str(e.value) == "different sparsevec dimensions 2 and 3"
    with pytest.raises(ValueError) as e:
        S.SparseVector(2, [2], [1.0])
    assert str(e.value) == "sparsevec index out of bounds"
    t = S.SparseTable(5)
    with pytest.raises(ValueError):
        t.exact_topk(O.L2, [S.SparseVector(4, [0], [1.0])], 1)
    # unsorted CSR handed straight to the C ABI is refused, not mis-scored
    bad = S.SparseRows(5, [0, 2], [3, 1], [1.0, 2.0])
    with pytest.raises(pv.VecB200Error) as e2:
        S.l2_distance(S.SparseVector(5, [0], [1.0]), bad)
    assert "ascending" in str(e2.value)
