"""Error behaviour of the C ABI on a GPU box: bad arguments are status codes + messages (the glue turns
them into ereport(ERROR)), never crashes, and the library keeps working afterwards."""
import ctypes as C

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pv():
    import pgvector_b200 as pv
    pv.init(0)
    return pv


def test_metric_must_fit_the_type(pv):
    with pytest.raises(pv.VecB200Error) as e:
        pv.distance_batch(O.VECTOR, O.HAMMING, np.zeros(4, np.float32), np.zeros((2, 4), np.float32))
    assert e.value.code == -1
    with pytest.raises(pv.VecB200Error):
        pv.distance_batch(O.BIT, O.L2, np.zeros(2, np.uint8), np.zeros((2, 2), np.uint8), dim=16)
    # and the library still works
    assert pv.l2_distance(np.zeros(4, np.float32), np.ones((1, 4), np.float32))[0] == 2.0


def test_dimension_mismatch_is_the_reference_error_text(pv):
    with pytest.raises(ValueError, match="different vector dimensions 2 and 3"):
        pv.l2_distance(np.zeros(2, np.float32), np.zeros((1, 3), np.float32))
    with pytest.raises(ValueError, match="different bit lengths 3 and 2"):
        pv.hamming_distance(np.zeros(1, np.uint8), np.zeros((1, 1), np.uint8), dim=2, q_dim=3)


def test_ivfflat_rejects_unsupported_opclasses_and_ranges(pv):
    with pytest.raises(ValueError):
        pv.IvfflatIndex("vector_l1_ops", 8, 4)          # l1 is hnsw only (sql/vector.sql:443-446)
    with pytest.raises(pv.VecB200Error):
        pv.IvfflatIndex("vector_l2_ops", 8, 0)          # lists 1..32768 (src/ivfflat.h:56-57)
    with pytest.raises(pv.VecB200Error):
        pv.IvfflatIndex("vector_l2_ops", 8, 40000)
    ix = pv.IvfflatIndex("vector_l2_ops", 8, 4)
    with pytest.raises(pv.VecB200Error) as e:
        ix.search(np.zeros((1, 8), np.float32), k=1, probes=1)   # not loaded
    assert "not loaded" in str(e.value)


def test_hnsw_parameter_ranges(pv):
    with pytest.raises(pv.VecB200Error):
        pv.HnswIndex("vector_l2_ops", 8, m=1)           # m 2..100 (src/hnsw.h:54-56)
    gi = pv.HnswIndex("vector_l2_ops", 3)
    nbr0 = np.full((1, 32), -1, np.int32)
    gi.load(np.array([[1, 2, 3]], np.float32), np.zeros(1, np.int32), nbr0, np.full(1, -1, np.int64), np.zeros((0, 16), np.int32), 0)
    with pytest.raises(pv.VecB200Error):
        gi.search(np.zeros((1, 3), np.float32), k=5, ef_search=1001)   # ef_search 1..1000 (src/hnsw.h:60-62)
    with pytest.raises(pv.VecB200Error):
        gi.search(np.zeros((1, 3), np.float32), k=50, ef_search=10)    # k <= ef


def test_kmeans_and_assign_argument_checks(pv):
    t = pv.Table(O.VECTOR, 4).append(np.random.default_rng(0).standard_normal((100, 4)).astype(np.float32))
    with pytest.raises(pv.VecB200Error):
        pv.kmeans(t, O.HAMMING, np.zeros((2, 4), np.float32))          # Hamming needs bit rows
    with pytest.raises(pv.VecB200Error):
        pv.assign(t, O.COSINE, np.zeros((2, 4), np.float32))           # proc 1 only
    out = pv.assign(t, O.L2_SQUARED, np.zeros((1, 4), np.float32))
    assert np.all(out == 0)


def test_empty_inputs_are_fine(pv):
    assert pv.distance_batch(O.VECTOR, O.L2, np.zeros(3, np.float32), np.zeros((0, 3), np.float32)).shape == (0,)
    t = pv.Table(O.VECTOR, 3)
    ids, dist = t.exact_topk(O.L2, np.zeros((2, 3), np.float32), 4)
    assert np.all(ids == -1) and np.all(np.isinf(dist))
    ix = pv.IvfflatIndex("vector_l2_ops", 3, 2).load(np.zeros((2, 3), np.float32), np.zeros(3, np.int64), np.zeros((0, 3), np.float32))
    ids, dist = ix.search(np.ones((3, 3), np.float32), k=2, probes=2)
    assert np.all(ids == -1)
