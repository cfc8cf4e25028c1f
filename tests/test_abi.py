"""CPU checks of the drop-in boundary: libvecb200.so loads and exports every symbol that
include/vecb200.h declares; constants agree with the oracle; the product never touches oracle/."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vecb200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if n != "vb_allreduce_fn"))


def test_header_declares_the_expected_surface():
    fns = declared_functions()
    for must in ("vb_init", "vb_last_error", "vb_distance_batch", "vb_exact_topk", "vb_ivf_scan_lists",
                 "vb_ivf_scan_items", "vb_ivf_search", "vb_kmeans", "vb_assign", "vb_hnsw_search"):
        assert must in fns
    assert len(fns) >= 30


def test_library_exports_every_declared_symbol():
    from pgvector_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [f for f in declared_functions() if not hasattr(lib, f)]
    assert not missing, missing
    assert lib.vb_abi_version() == 1
    # the ctypes table binds exactly the declared surface
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_no_device_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is visible")
    import numpy as np
    import pgvector_b200 as pv
    with pytest.raises(pv.VecB200Error) as e:
        pv.l2_distance(np.zeros(3, np.float32), np.zeros((2, 3), np.float32))
    assert e.value.code == -2
    assert "no CPU fallback" in str(e.value)


def test_metric_and_type_codes_match_the_oracle():
    import oracle as O
    import pgvector_b200 as pv
    hdr = open(HEADER).read()
    for name in ("VECTOR", "HALFVEC", "BIT", "L2_SQUARED", "NEG_IP", "COSINE", "L1", "HAMMING", "JACCARD", "L2", "IP", "SPHERICAL"):
        m = re.search(rf"#define VB_{name} (\d+)", hdr)
        assert m, name
        assert int(m.group(1)) == getattr(O, name) == getattr(pv, name)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pgvector_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cpp")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "pgv_oracle.h" not in text, f
