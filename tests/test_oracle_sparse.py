"""CPU tests: the sparsevec oracle (oracle/pgv_sparse.c) against the reference's known-answer outputs
(tests/golden/sparsevec_kat.json, transcribed from test/expected/sparsevec.out), the tiny index orderings of
test/expected/hnsw_sparsevec.out, and the dense oracle on densified vectors."""
import math

import numpy as np
import pytest

import oracle as O
from pgvector_b200.sparsevec import SparseRows, SparseVector
from tests.util import load_golden

KAT = load_golden("sparsevec_kat.json")["cases"]
ORDERINGS = load_golden("sparsevec_orderings.json")["blocks"]
METRIC = {"l2_distance": O.L2, "inner_product": O.IP, "negative_inner_product": O.NEG_IP, "cosine_distance": O.COSINE,
          "l1_distance": O.L1}
OPS = {"<->": O.L2, "<#>": O.NEG_IP, "<=>": O.COSINE, "<+>": O.L1}


def expect_float(text):
    return math.nan if text == "NaN" else float(text.replace("Infinity", "inf"))


def sv(v):
    return (v.indices, v.values)


@pytest.mark.parametrize("case", KAT, ids=[c["source"].split("/")[-1] for c in KAT])
def test_known_answer(case):
    fn = case["fn"]
    args = [SparseVector.from_text(a) for a in case["args"]]
    if fn in METRIC:
        a, b = args
        if case["error"]:
            assert case["error"] == f"different sparsevec dimensions {a.dim} and {b.dim}"   # CheckDims, src/sparsevec.c:44-51
            return
        got = O.sparse_distance(METRIC[fn], sv(a), sv(b))
        want = expect_float(case["expected"])
        assert (math.isnan(got) and math.isnan(want)) or got == want, (case, got)
    elif fn == "l2_norm":
        got = O.sparse_l2_norm(sv(args[0]))
        if case["real"]:
            assert np.float32(got) == np.float32(expect_float(case["expected"]))
        else:
            assert got == expect_float(case["expected"])
    elif fn == "l2_normalize":
        gi, gv = O.sparse_l2_normalize(sv(args[0]))
        want = SparseVector.from_text(case["expected"])
        assert np.array_equal(gi, want.indices) and np.array_equal(gv, want.values), (gi, gv, want)
    else:
        pytest.fail(f"unhandled {fn}")


def test_kat_coverage():
    fns = {c["fn"] for c in KAT}
    assert {"l2_distance", "inner_product", "cosine_distance", "l1_distance", "l2_norm", "l2_normalize"} <= fns
    assert len(KAT) >= 45


def random_sparse(rng, dim, nnz):
    idx = np.sort(rng.choice(dim, size=nnz, replace=False)) if dim < 10 * max(nnz, 1) else np.unique(rng.integers(0, dim, size=nnz))
    return SparseVector(dim, idx, rng.standard_normal(idx.size).astype(np.float32))


def test_sparse_equals_dense_arithmetic_on_small_integers():
    """every term and its order are those of the dense loops when values are small integers (exact in fp32):
    sparsevec and vector functions agree exactly (test/t/034_distance_functions.pl does the same across types)"""
    rng = np.random.default_rng(5)
    for _ in range(200):
        dim = int(rng.integers(1, 40))
        a = SparseVector.from_dense(rng.integers(-3, 4, size=dim).astype(np.float32))
        b = SparseVector.from_dense(rng.integers(-3, 4, size=dim).astype(np.float32))
        for m in (O.L2_SQUARED, O.L2, O.IP, O.NEG_IP, O.L1, O.COSINE):
            s = O.sparse_distance(m, sv(a), sv(b))
            d = O.distance(O.VECTOR, m, a.to_dense(), b.to_dense())
            assert s == d or (math.isnan(s) and math.isnan(d)), (m, a, b, s, d)


def test_fp32_sums_within_tolerance_of_truth():
    rng = np.random.default_rng(6)
    for dim, nnz in ((50, 20), (10_000, 300), (1_000_000_000, 1000), (1_000_000, 16_000)):
        a, b = random_sparse(rng, dim, nnz), random_sparse(rng, dim, nnz)
        # force overlap
        b = SparseVector(dim, np.concatenate([b.indices[: b.nnz // 2], a.indices[: a.nnz // 2]]),
                         np.concatenate([b.values[: b.nnz // 2], rng.standard_normal(a.nnz // 2).astype(np.float32)])) \
            if not set(a.indices[: a.nnz // 2]) & set(b.indices[: b.nnz // 2]) else b
        for m in (O.L2_SQUARED, O.L2, O.L1, O.COSINE):
            t = O.sparse_distance(m, sv(a), sv(b), f64=True)
            assert abs(O.sparse_distance(m, sv(a), sv(b)) - t) <= 1e-5 * max(abs(t), 1e-30)
        t = O.sparse_distance(O.IP, sv(a), sv(b), f64=True)
        scale = float(np.abs(a.values).max() * np.abs(b.values).max()) * max(a.nnz, 1)
        assert abs(O.sparse_distance(O.IP, sv(a), sv(b)) - t) <= 1e-5 * scale


def test_batch_equals_pairs_and_argument_order():
    rng = np.random.default_rng(7)
    dim = 500
    rows = [random_sparse(rng, dim, int(rng.integers(0, 60))) for _ in range(64)]
    q = random_sparse(rng, dim, 40)
    R = SparseRows.from_vectors(rows, dim)
    for m in (O.L2, O.NEG_IP, O.COSINE, O.L1):
        got = O.sparse_distance_batch(m, sv(q), R.row_off, R.idx, R.val)
        for r, v in enumerate(rows):
            w = O.sparse_distance(m, sv(v), sv(q))
            assert got[r] == w or (math.isnan(got[r]) and math.isnan(w))


@pytest.mark.parametrize("block", ORDERINGS, ids=[b["index"]["opclass"] for b in ORDERINGS])
def test_tiny_index_orderings(block):
    """test/expected/hnsw_sparsevec.out: on four rows the index returns the exact order"""
    vals = [v for grp in block["rows"] for v in grp["values"] if v is not None]
    rows = [SparseVector.from_text(v) for v in vals]
    for qd in block["queries"]:
        q = SparseVector.from_text(qd["query"])
        d = np.array([O.sparse_distance(OPS[qd["op"]], sv(r), sv(q)) for r in rows])
        order = [i for i in np.argsort(d, kind="stable") if not math.isnan(d[i])]   # the cosine opclass does not index zero vectors
        assert [vals[i] for i in order] == qd["expected"]
