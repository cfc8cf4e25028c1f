"""CPU tests of the sparsevec host mirror (pgvector_b200/sparsevec.py): text I/O, validation with the reference's error
texts (src/sparsevec.c:66-150, 215-395), CSR packing.  No device call is made here."""
import numpy as np
import pytest

from pgvector_b200.sparsevec import SPARSEVEC_MAX_NNZ, SparseRows, SparseVector
from tests.util import load_golden


def test_text_round_trip_of_every_reference_literal():
    """every literal of test/expected/sparsevec.out that the known-answer file uses parses and prints back canonically"""
    seen = set()
    for case in load_golden("sparsevec_kat.json")["cases"]:
        for lit in case["args"] + ([case["expected"]] if case["fn"] == "l2_normalize" and case["expected"] else []):
            if lit in seen:
                continue
            seen.add(lit)
            v = SparseVector.from_text(lit)
            again = SparseVector.from_text(v.to_text())
            assert again.dim == v.dim and np.array_equal(again.indices, v.indices) and np.array_equal(again.values, v.values)
    assert len(seen) > 30


def test_canonical_text_of_simple_values():
    assert SparseVector.from_text("{1:1,3:2.5}/5").to_text() == "{1:1,3:2.5}/5"
    assert SparseVector.from_text(" { 3:2 , 1:1 } / 4 ".replace(" ", "")).to_text() == "{1:1,3:2}/4"      # sorted by index
    assert SparseVector.from_text("{1:0,2:3}/2").to_text() == "{2:3}/2"                                  # zeros are not stored
    assert SparseVector.from_text("{}/3").nnz == 0


@pytest.mark.parametrize("bad,text", [
    ("{1:1}/0", "sparsevec must have at least 1 dimension"),
    ("{3:1}/2", "sparsevec index out of bounds"),
    ("{0:1}/2", "sparsevec index out of bounds"),
    ("{1:1,1:2}/2", "sparsevec indices must not contain duplicates"),
])
def test_reference_error_texts(bad, text):
    with pytest.raises(ValueError) as e:
        SparseVector.from_text(bad)
    assert str(e.value) == text


@pytest.mark.parametrize("bad", ["", "1:1/2", "{1:1}", "{1}/2", "{1:x}/2", "{1:1}/x"])
def test_syntax_errors(bad):
    with pytest.raises(ValueError) as e:
        SparseVector.from_text(bad)
    assert str(e.value) == f'invalid input syntax for type sparsevec: "{bad}"'


def test_value_checks():
    with pytest.raises(ValueError, match="NaN not allowed in sparsevec"):
        SparseVector(3, [0], [np.nan])
    with pytest.raises(ValueError, match="infinite value not allowed in sparsevec"):
        SparseVector(3, [0], [np.inf])
    with pytest.raises(ValueError, match=f"sparsevec cannot have more than {SPARSEVEC_MAX_NNZ} non-zero elements"):
        SparseVector(100000, np.arange(SPARSEVEC_MAX_NNZ + 1), np.ones(SPARSEVEC_MAX_NNZ + 1))


def test_csr_packing_and_dense_round_trip():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((7, 40)).astype(np.float32)
    x[rng.random(x.shape) < 0.7] = 0
    x[3] = 0
    R = SparseRows.from_dense(x)
    assert R.n == 7 and R.dim == 40 and R.row_off[0] == 0 and R.row_off[-1] == np.count_nonzero(x)
    for r in range(7):
        assert np.array_equal(R.row(r).to_dense(), x[r])
    R2 = SparseRows.from_vectors([SparseVector.from_dense(row) for row in x])
    assert np.array_equal(R.row_off, R2.row_off) and np.array_equal(R.idx, R2.idx) and np.array_equal(R.val, R2.val)
    with pytest.raises(ValueError, match="expected 40 dimensions, not 3"):
        SparseRows.from_vectors([SparseVector.from_dense(x[0]), SparseVector(3, [0], [1.0])])
