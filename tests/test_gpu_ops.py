"""GPU parity of the batched row transforms (norm, l2_normalize, binary_quantize) vs the reference's
known answers and the oracle."""
import numpy as np
import pytest

import oracle as O
from tests.util import f32_to_half_bits, half_bits_to_f32, load_golden, parse_vector

pytestmark = pytest.mark.gpu
ELEM = {"vector": O.VECTOR, "halfvec": O.HALFVEC}
KAT = [c for c in load_golden("distance_kat.json")["cases"] if c["fn"] in ("vector_norm", "l2_norm", "l2_normalize", "binary_quantize")]


@pytest.fixture(scope="module")
def pv():
    import pgvector_b200 as pv
    pv.init(0)
    return pv


@pytest.mark.parametrize("case", KAT, ids=[c["source"].split("/")[-1] for c in KAT])
def test_known_answers(pv, case):
    elem = ELEM[case["type"]]
    (a, _), = [parse_vector(x, elem) for x in case["args"]]
    if case["fn"] in ("vector_norm", "l2_norm"):
        got = pv.vector_norm(a, elem)
        if case["expected"] is None:
            return
        want = float(case["expected"])
        if case["real"]:
            assert np.float32(got) == np.float32(want)
        else:
            assert got == want
    elif case["fn"] == "l2_normalize":
        if case["error"]:
            with pytest.raises(OverflowError):
                pv.l2_normalize(a, elem)
            return
        got = pv.l2_normalize(a, elem)
        want, _ = parse_vector(case["expected"], elem)
        assert np.array_equal(got, want)
    else:
        got = pv.binary_quantize(a, elem)
        want, _ = parse_vector(case["expected"], O.BIT)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("elem,dim", [(O.VECTOR, 3), (O.VECTOR, 1536), (O.HALFVEC, 9), (O.HALFVEC, 768)])
def test_random_rows_match_oracle(pv, elem, dim):
    rng = np.random.default_rng(dim)
    x = rng.standard_normal((500, dim)).astype(np.float32)
    x[7] = 0          # zero row stays zero
    rows = f32_to_half_bits(x) if elem == O.HALFVEC else x
    norms = pv.vector_norm(rows, elem)
    want_n = np.array([O.norm(elem, r) for r in rows])
    assert np.allclose(norms, want_n, rtol=1e-14)
    got = pv.l2_normalize(rows, elem)
    want = O.l2_normalize(elem, rows)
    if elem == O.HALFVEC:
        g, w = half_bits_to_f32(got), half_bits_to_f32(want)
        assert np.mean(got == want) > 0.999 and np.allclose(g, w, rtol=1e-3, atol=1e-7)
    else:
        assert np.mean(got == want) > 0.999 and np.allclose(got, want, rtol=2e-7, atol=0)
    assert not got[7].any()
    assert np.array_equal(pv.binary_quantize(rows, elem), O.binary_quantize(elem, rows))


def test_vector_to_halfvec_cast(pv):
    """vector::halfvec (src/halfvec.c:540-555, Float4ToHalf src/halfutils.h:244-261): RNE, equal to the reference's own
    converter bit for bit, and the reference's error text when a finite value does not fit (test/expected/cast.out:268-269)"""
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.standard_normal(5000) * 100, [0.0, -0.0, 65504.0, -65504.0, 65519.0, 5.9604645e-08, 2.9802322e-08, 1e-10, np.inf, -np.inf],
                        rng.standard_normal(2000) * 1e-5]).astype(np.float32).reshape(-1, 2)
    got = pv.vector_to_halfvec(x)
    want = np.array([O.lib().pgv_float_to_half(float(v)) for v in x.ravel()], dtype=np.uint16).reshape(x.shape)
    assert np.array_equal(got, want)
    back = pv.halfvec_to_vector(got)
    assert np.array_equal(back, want.view(np.float16).astype(np.float32))
    with pytest.raises(ValueError, match='"65520" is out of range for type halfvec'):
        pv.vector_to_halfvec(np.array([[1.0, 2.0], [65520.0, -65520.0]], dtype=np.float32))
    with pytest.raises(ValueError, match=r'"-3e\+38" is out of range for type halfvec'):
        pv.vector_to_halfvec(np.array([1.0, -3e38, 7e4], dtype=np.float32))
    with pytest.raises(ValueError, match='"100000" is out of range for type halfvec'):
        pv.vector_to_halfvec(np.array([1e5], dtype=np.float32))
