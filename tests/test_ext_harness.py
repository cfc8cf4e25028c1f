"""The extension glue (pgvector_b200/ext/*.c: page packers, scan bodies, build bodies) RUN over synthesised index
pages.  The harness (tests/harness) compiles the glue against the reference's own headers and a functional stand-in
for the server (pgstub_runtime.c); the C ABI behind it is the oracle-backed mock on the CPU and the real
libvecb200.so on the GPU box.  What is checked: the packers recover exactly the arrays the pages were written from
(1-byte and 4-byte varlena headers, multi-page list / entry chains, empty lists, element renumbering, duplicate heap
TIDs, elements being deleted), a scan through VbGetScanLists / VbGetScanItems / VbNextItem (and the HNSW pair)
returns the oracle's heap TIDs, the image cache honours the cross-backend version stamp, and an ERROR inside a C ABI
call leaves neither device handles nor pinned buffers behind."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle as O
from tests.harness import build as hbuild
from tests.harness import pages as P
from tests.util import build_ivf_arrays, f32_to_half_bits, mixture

HAVE_MOCK, HAVE_REAL = hbuild.build()
PROC = {("vector", "l2"): 0, ("vector", "ip"): 1, ("vector", "l1"): 2, ("halfvec", "l2"): 3, ("halfvec", "ip"): 4, ("halfvec", "l1"): 5,
        ("bit", "hamming"): 6, ("bit", "jaccard"): 7, ("vector", "kmeans_l2"): 8, ("vector", "kmeans_spherical"): 9,
        ("halfvec", "kmeans_l2"): 10, ("bit", "kmeans_hamming"): 6}
ELEM = {"vector": O.VECTOR, "halfvec": O.HALFVEC, "bit": O.BIT}
_relid = [1000]


class Harness:
    def __init__(self, real):
        mock, realp = hbuild.paths()
        self.real = real
        self.lib = C.CDLL(realp if real else mock)
        L = self.lib
        L.h_open.restype = C.c_void_p
        L.h_open.argtypes = [C.c_void_p, C.c_uint32, C.c_uint, C.c_int, C.c_int, C.c_int]
        L.pgstub_last_error.restype = C.c_char_p
        L.h_ivf_scan.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.h_ivf_image.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.h_bump_version.argtypes = [C.c_void_p]
        L.h_ivf_invalidate.argtypes = [C.c_void_p]
        L.h_hnsw_invalidate.argtypes = [C.c_void_p]
        L.h_hnsw_scan.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.h_ivf_kmeans.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.h_ivf_assign.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.pgstub_pinned_buffers.restype = C.c_int
        L.pgstub_buffer_reads.restype = C.c_long
        L.h_hnsw_set_iterative.argtypes = [C.c_int, C.c_int]
        if not real:
            L.mock_live.restype = C.c_int
            L.mock_ivf_rows.restype = C.c_int64
            L.mock_ivf_rows.argtypes = [C.c_void_p] + [C.c_void_p] * 4
            L.mock_hnsw_graph.restype = C.c_int64
            L.mock_hnsw_graph.argtypes = [C.c_void_p] + [C.c_void_p] * 7

    def err(self):
        return self.lib.pgstub_last_error().decode()

    def open(self, image, dim, proc1, proc3=-1):
        buf = C.create_string_buffer(image, len(image))
        _relid[0] += 1
        rel = self.lib.h_open(buf, len(image) // P.BLCKSZ, _relid[0], dim, proc1, proc3)
        return rel, buf          # keep buf alive: the relation's pages

    def ivf_scan(self, rel, elem, q, probes, max_probes, max_items, short=0):
        out = np.empty(max(max_items, 1), dtype=np.int64)
        n, nb = C.c_int64(), C.c_int64()
        qp = None if q is None else np.ascontiguousarray(q).ctypes.data_as(C.c_void_p)
        rc = self.lib.h_ivf_scan(rel, elem, qp, short, probes, max_probes, max_items, out.ctypes.data_as(C.c_void_p), C.byref(n), C.byref(nb))
        if rc != 0:
            raise RuntimeError(self.err())
        return out[:n.value], nb.value

    def ivf_image(self, rel, max_lists=4096):
        lists, reads, ix = C.c_int(), C.c_long(), C.c_void_p()
        sp = np.empty(max_lists, dtype=np.uint32)
        rc = self.lib.h_ivf_image(rel, C.byref(lists), sp.ctypes.data_as(C.c_void_p), max_lists, C.byref(reads), C.byref(ix))
        if rc != 0:
            raise RuntimeError(self.err())
        return lists.value, sp[:lists.value], reads.value, ix

    def mock_ivf(self, ix, elem, dim, lists):
        c, o, r, i = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = self.lib.mock_ivf_rows(ix, C.byref(c), C.byref(o), C.byref(r), C.byref(i))
        rb = P.row_bytes(elem, dim)
        grab = lambda p, nbytes: np.frombuffer(C.string_at(p.value, nbytes), dtype=np.uint8).copy() if nbytes else np.zeros(0, np.uint8)
        return (grab(c, rb * lists).reshape(lists, rb), np.frombuffer(C.string_at(o.value, 8 * (lists + 1)), dtype=np.int64).copy(),
                grab(r, rb * n).reshape(n, rb), np.frombuffer(C.string_at(i.value, 8 * n), dtype=np.int64).copy() if n else np.zeros(0, np.int64))

    def hnsw_scan(self, rel, elem, q, ef, max_items):
        out = np.empty(max(max_items, 1), dtype=np.int64)
        n, tuples = C.c_int64(), C.c_int64()
        qp = None if q is None else np.ascontiguousarray(q).ctypes.data_as(C.c_void_p)
        rc = self.lib.h_hnsw_scan(rel, elem, qp, ef, max_items, out.ctypes.data_as(C.c_void_p), C.byref(n), C.byref(tuples))
        if rc != 0:
            raise RuntimeError(self.err())
        return (None if n.value == -2 else out[:n.value]), tuples.value


def _params(fn):
    cpu = pytest.param(False, id="mock-abi", marks=pytest.mark.skipif(not HAVE_MOCK, reason="harness not built (no reference headers)"))
    gpu = pytest.param(True, id="libvecb200", marks=[pytest.mark.gpu, pytest.mark.skipif(not HAVE_REAL, reason="harness not built")])
    return pytest.mark.parametrize("real", [cpu, gpu])(fn)


def tids_of(rows_numbers):
    return np.array([P.tid_id(*P.heap_tid_of(int(r))) for r in rows_numbers], dtype=np.int64)


def ivf_case(typ, metric, n, dim, lists, seed):
    elem = ELEM[typ]
    x, c = mixture(n, dim, lists, seed=seed)
    if elem == O.BIT:
        rows, centers = O.binary_quantize(O.VECTOR, x), O.binary_quantize(O.VECTOR, c)
    elif elem == O.HALFVEC:
        rows, centers = f32_to_half_bits(x), f32_to_half_bits(c)
    else:
        rows, centers = x, c
    m = {"l2": O.L2_SQUARED, "ip": O.NEG_IP, "hamming": O.HAMMING}[metric]
    assign = O.ivf_assign(elem, m, rows, centers, threads=4, dim=dim)
    assign[assign == 1] = 0                     # list 1 is empty on purpose
    grouped, order, offsets = build_ivf_arrays(rows, assign, lists)
    return elem, m, rows, centers, grouped, order, offsets


@_params
@pytest.mark.parametrize("typ,metric,n,dim,lists,per_page", [("vector", "l2", 3000, 3, 40, 7), ("vector", "l2", 2500, 40, 300, None),
                                                          ("halfvec", "l2", 2000, 20, 12, None), ("halfvec", "ip", 1500, 200, 9, 5),
                                                          ("bit", "hamming", 2500, 52, 10, None), ("bit", "hamming", 1200, 1024, 8, 11)])
def test_ivfflat_packer_and_scan(real, typ, metric, n, dim, lists, per_page):
    H = Harness(real)
    O.ivf_set_tie_mode(True)
    try:
        elem, m, rows, centers, grouped, order, offsets = ivf_case(typ, metric, n, dim, lists, seed=n + dim)
        image, info = P.ivfflat_image(elem, dim, centers, offsets, grouped, order, entries_per_page=per_page)
        rel, keep = H.open(image, dim, PROC[(typ, metric)])
        nl, start_pages, reads, ix = H.ivf_image(rel)
        assert nl == lists and list(start_pages) == info["start_pages"]
        assert H.lib.pgstub_pinned_buffers() == 0
        if not real:
            c, o, r, i = H.mock_ivf(ix, elem, dim, lists)
            assert np.array_equal(o, offsets)
            assert np.array_equal(c, np.ascontiguousarray(centers).view(np.uint8).reshape(lists, -1))
            assert np.array_equal(r, np.ascontiguousarray(grouped).view(np.uint8).reshape(n, -1))
            assert np.array_equal(i, tids_of(order))
        oix = O.Ivf(elem, m, centers, offsets, grouped, tids_of(order), dim=dim)
        rng = np.random.default_rng(1)
        for qi in rng.integers(0, n, 12):
            q = rows[qi]
            got, batches = H.ivf_scan(rel, elem, q, probes=3, max_probes=3, max_items=10 ** 6)
            want, wd, _ = oix.search(q, 3, 0)
            assert batches == 1 and len(got) == len(want)
            if elem == O.BIT:
                assert np.array_equal(got, want)
            else:
                assert (got == want).mean() > 0.99 and sorted(got) == sorted(want)
        # the second scan of the same pages re-used the image: one meta-page read per scan, no repack
        _, _, reads2, _ = H.ivf_image(rel)
        assert reads2 == 1
        # iterative scan (src/ivfscan.c:400-406): batches of `probes` lists until max_probes
        q = rows[5]
        got, batches = H.ivf_scan(rel, elem, q, probes=2, max_probes=6, max_items=10 ** 6)
        wl, _ = oix.scan_lists(q, 6)
        want = np.concatenate([oix_items(oix, q, wl[i:i + 2]) for i in range(0, len(wl), 2)])
        assert batches == (len(wl) + 1) // 2
        assert len(got) == len(want) and (elem != O.BIT or np.array_equal(got, want))
        # LIMIT: the scan stops pulling
        got10, _ = H.ivf_scan(rel, elem, q, probes=3, max_probes=3, max_items=10)
        assert len(got10) == 10
        # NULL query: every distance 0, every row of the probed lists comes back (src/ivfscan.c:207-211)
        got0, _ = H.ivf_scan(rel, elem, None, probes=2, max_probes=2, max_items=10 ** 6)
        assert len(got0) > 0 and len(set(got0.tolist())) == len(got0)
    finally:
        O.ivf_set_tie_mode(False)


def oix_items(oix, q, lists):
    L = O.lib()
    lists = np.ascontiguousarray(lists, dtype=np.int32)
    total = int(sum(oix.offsets[l + 1] - oix.offsets[l] for l in lists))
    ids = np.empty(max(total, 1), dtype=np.int64)
    dist = np.empty(max(total, 1), dtype=np.float64)
    qq = np.ascontiguousarray(q)
    L.pgv_ivf_scan_items(C.byref(oix.c), qq.ctypes.data_as(C.c_void_p), lists.ctypes.data_as(C.c_void_p), len(lists), total,
                         ids.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p))
    return ids[:total]


@pytest.mark.skipif(not HAVE_MOCK, reason="harness not built (no reference headers)")
def test_short_varlena_queries_and_invalidation_stamp():
    H = Harness(False)
    elem, m, rows, centers, grouped, order, offsets = ivf_case("vector", "l2", 1500, 3, 10, seed=5)
    image, info = P.ivfflat_image(elem, 3, centers, offsets, grouped, order)
    rel, keep = H.open(image, 3, PROC[("vector", "l2")])
    a, _ = H.ivf_scan(rel, elem, rows[7], 2, 2, 50, short=0)
    b, _ = H.ivf_scan(rel, elem, rows[7], 2, 2, 50, short=1)     # the query datum itself arrives with a 1-byte header
    assert np.array_equal(a, b)
    _, _, reads, ix1 = H.ivf_image(rel)
    assert reads == 1                                             # cached: only the version stamp was read
    live = H.lib.mock_live()
    assert H.lib.h_bump_version(rel) == 0                          # what aminsert / ambulkdelete do in ANY backend
    _, _, reads, ix2 = H.ivf_image(rel)
    assert reads > 5 and H.lib.mock_live() == live                 # repacked, the old image was released
    H.lib.h_ivf_invalidate(rel)
    assert H.lib.mock_live() == live - 1
    _, _, reads, _ = H.ivf_image(rel)
    assert reads > 5 and H.lib.mock_live() == live
    H.lib.h_ivf_invalidate(rel)


@pytest.mark.skipif(not HAVE_MOCK, reason="harness not built (no reference headers)")
def test_error_in_the_c_abi_leaves_nothing_behind():
    H = Harness(False)
    elem, m, rows, centers, grouped, order, offsets = ivf_case("vector", "l2", 800, 8, 6, seed=9)
    image, _ = P.ivfflat_image(elem, 8, centers, offsets, grouped, order)
    rel, keep = H.open(image, 8, PROC[("vector", "l2")])
    before = H.lib.mock_live()
    H.lib.mock_fail_next(1)
    with pytest.raises(RuntimeError, match="vecb200: mock: injected load failure"):
        H.ivf_scan(rel, elem, rows[0], 2, 2, 10)
    assert H.lib.mock_live() == before and H.lib.pgstub_pinned_buffers() == 0
    got, _ = H.ivf_scan(rel, elem, rows[0], 2, 2, 10)             # the next scan packs again and works
    assert len(got) == 10
    # the build body: PG_FINALLY frees the sample table when a call inside PG_TRY raises
    rel2, keep2 = H.open(image, 8, PROC[("vector", "l2")], PROC[("vector", "kmeans_l2")])
    live = H.lib.mock_live()
    H.lib.mock_fail_next(1)
    cen = np.empty((5, 8), dtype=np.float32)
    used = C.c_int()
    rc = H.lib.h_ivf_kmeans(rel2, elem, 8, rows.ctypes.data_as(C.c_void_p), 400, 5, cen.ctypes.data_as(C.c_void_p), C.byref(used))
    assert rc == -1 and "injected" in H.err() and H.lib.mock_live() == live
    H.lib.h_ivf_invalidate(rel)


@_params
def test_ivfflat_build_bodies(real):
    """VbIvfflatKmeans (k-means++ + k-means on the sampled rows) and the assign batch of the build callback"""
    H = Harness(real)
    n, dim, lists = 3000, 24, 16
    x, _ = mixture(n, dim, lists, seed=77)
    image, _ = P.ivfflat_image(O.VECTOR, dim, x[:1], np.array([0, 0]), x[:0], [])
    rel, keep = H.open(image, dim, PROC[("vector", "l2")], PROC[("vector", "kmeans_l2")])
    cen = np.zeros((lists, dim), dtype=np.float32)
    used = C.c_int()
    assert H.lib.h_ivf_kmeans(rel, O.VECTOR, dim, x.ctypes.data_as(C.c_void_p), n, lists, cen.ctypes.data_as(C.c_void_p), C.byref(used)) == 0, H.err()
    assert used.value == 1 and np.isfinite(cen).all()
    # fewer samples than lists: the reference path must run (src/ivfkmeans.c:110-133)
    assert H.lib.h_ivf_kmeans(rel, O.VECTOR, dim, x.ctypes.data_as(C.c_void_p), 5, lists, cen.copy().ctypes.data_as(C.c_void_p), C.byref(used)) == 0
    assert used.value == 0
    out_l = np.empty(n, dtype=np.int32)
    out_t = np.empty(n, dtype=np.int64)
    assert H.lib.h_ivf_assign(rel, O.VECTOR, dim, x.ctypes.data_as(C.c_void_p), n, cen.ctypes.data_as(C.c_void_p), lists, 700,
                              out_l.ctypes.data_as(C.c_void_p), out_t.ctypes.data_as(C.c_void_p)) == 0, H.err()
    want = O.ivf_assign(O.VECTOR, O.L2_SQUARED, x, cen, threads=4)
    assert (out_l == want).mean() > 0.9995
    assert np.array_equal(out_t, np.array([((i // 100) << 16) | (i % 100 + 1) for i in range(n)]))


def hnsw_case(typ, metric, n, dim, m, seed, dups=0):
    elem = ELEM[typ]
    x, _ = mixture(n, dim, 20, seed=seed)
    if elem == O.BIT:
        rows = O.binary_quantize(O.VECTOR, x)
    elif elem == O.HALFVEC:
        rows = f32_to_half_bits(x)
    else:
        rows = x
    if dups:
        rows = np.concatenate([rows, rows[:dups]])
    mt = {"l2": O.L2_SQUARED, "ip": O.NEG_IP, "l1": O.L1, "hamming": O.HAMMING, "jaccard": O.JACCARD}[metric]
    og = O.Hnsw(elem, mt, rows, m=m, ef_construction=40, seed=seed, dim=dim)
    return elem, mt, rows, og, og.export()


@_params
@pytest.mark.parametrize("typ,metric,n,dim,m,dups", [("vector", "l2", 2500, 16, 8, 40), ("halfvec", "l2", 1500, 40, 16, 0),
                                                    ("bit", "hamming", 2000, 256, 12, 0), ("vector", "l1", 1200, 6, 5, 0)])
def test_hnsw_packer_and_scan(real, typ, metric, n, dim, m, dups):
    H = Harness(real)
    elem, mt, rows, og, g = hnsw_case(typ, metric, n, dim, m, seed=n + dim, dups=dups)
    ne = len(g["levels"])
    erows = rows[g["elem_row"]]
    heaptids = [list(g["heaptids"][e][:g["n_heaptids"][e]]) for e in range(ne)]
    if dups:
        assert sum(len(h) for h in heaptids) == len(rows) and max(len(h) for h in heaptids) >= 2
    # CreateGraphPages writes the newest element first, so the packer renumbers the elements (by page position).  The
    # search breaks distance ties by element number: for the constantly-tying Hamming metric keep the numbering (write
    # oldest first) so the walk is comparable step by step; the float cases run on the reversed numbering.
    image, info = P.hnsw_image(elem, dim, m, erows, g["levels"], g["nbr0"], g["upper_off"], g["upper"], g["entry"], heaptids=heaptids,
                               write_order=range(ne) if elem == O.BIT else None)
    rel, keep = H.open(image, dim, PROC[(typ, metric)])
    rng = np.random.default_rng(2)
    for qi in rng.integers(0, len(rows), 25):
        q = rows[qi]
        got, tuples = H.hnsw_scan(rel, elem, q, ef=40, max_items=10 ** 6)
        wi, wd, wnd = og.search(q, 40, ties=O.TIES_TOTAL)
        # nearest element first, its heap TIDs last-added first (src/hnswscan.c:293-311)
        want = [P.tid_id(*P.heap_tid_of(int(r))) for e in wi for r in reversed(heaptids[e])]
        assert tuples == wnd
        if elem == O.BIT:
            assert list(got) == want
        else:
            assert len(got) == len(want) and (np.array(got) == np.array(want)).mean() > 0.97
    assert H.lib.pgstub_pinned_buffers() == 0
    # NULL query: served by the reference loop
    got, _ = H.hnsw_scan(rel, elem, None, ef=40, max_items=10)
    assert got is None
    H.lib.h_hnsw_invalidate(rel)


@_params
@pytest.mark.parametrize("typ,metric,n,dim,m,dups", [("vector", "l2", 2500, 16, 8, 40), ("halfvec", "ip", 1500, 40, 16, 0)])
def test_hnsw_iterative_scan_through_the_glue(real, typ, metric, n, dim, m, dups):
    """hnsw.iterative_scan (src/hnswscan.c:62-87, 236-262) through VbHnswGetScanItems / VbHnswNextItem: relaxed order
    returns the oracle's sequence (batch after batch, then the drain past hnsw.max_scan_tuples), strict order drops the
    elements nearer than one already returned (:316-322), a LIMIT stops pulling, and hnswendscan frees the handle."""
    H = Harness(real)
    elem, mt, rows, og, g = hnsw_case(typ, metric, n, dim, m, seed=n + dim, dups=dups)
    ne = len(g["levels"])
    erows = rows[g["elem_row"]]
    heaptids = [list(g["heaptids"][e][:g["n_heaptids"][e]]) for e in range(ne)]
    image, info = P.hnsw_image(elem, dim, m, erows, g["levels"], g["nbr0"], g["upper_off"], g["upper"], g["entry"], heaptids=heaptids,
                               write_order=range(ne))
    rel, keep = H.open(image, dim, PROC[(typ, metric)])
    tids = lambda elems: [P.tid_id(*P.heap_tid_of(int(r))) for e in elems for r in reversed(heaptids[e])]
    H.hnsw_scan(rel, elem, rows[0], ef=20, max_items=5)                          # (packs the image: one library handle)
    live0 = H.lib.mock_live() if not real else 0
    rng = np.random.default_rng(4)
    try:
        for qi in rng.integers(0, len(rows), 6):
            q = rows[qi]
            for max_tuples in (10 ** 6, 600):
                H.lib.h_hnsw_set_iterative(1, max_tuples)                      # relaxed_order
                got, tuples = H.hnsw_scan(rel, elem, q, ef=20, max_items=10 ** 6)
                wi, wd, wb, wt = og.iter_scan(q, 20, max_scan_tuples=max_tuples, ties=O.TIES_TOTAL)
                want = tids(wi)
                assert len(got) == len(set(got.tolist()))
                if real:
                    assert abs(len(got) - len(want)) <= 40 and len(set(got.tolist()) & set(want)) >= 0.97 * len(want)
                else:
                    assert list(got) == want and tuples == wt
                if max_tuples == 10 ** 6:
                    assert len(got) >= 0.9 * len(rows)                         # every reachable row comes back once
                got10, _ = H.hnsw_scan(rel, elem, q, ef=20, max_items=55)      # LIMIT 55: three batches are enough
                assert len(got10) == 55 and (real or list(got10) == want[:55])
            H.lib.h_hnsw_set_iterative(2, 600)                                 # strict_order
            got, _ = H.hnsw_scan(rel, elem, q, ef=20, max_items=10 ** 6)
            wi, wd, wb, wt = og.iter_scan(q, 20, max_scan_tuples=600, ties=O.TIES_TOTAL)
            keep_e, prev = [], -np.inf
            for e, d in zip(wi, wd):
                if d >= prev:
                    keep_e.append(int(e))
                    prev = d
            if not real:
                assert list(got) == tids(keep_e)
            assert 20 <= len(got) <= len(tids(wi))
    finally:
        H.lib.h_hnsw_set_iterative(0, 20000)
    if not real:
        assert H.lib.mock_live() == live0                                      # every scan handle was released
    assert H.lib.pgstub_pinned_buffers() == 0
    H.lib.h_hnsw_invalidate(rel)


@pytest.mark.skipif(not HAVE_MOCK, reason="harness not built (no reference headers)")
def test_hnsw_elements_being_deleted_and_unmapped_entry_point():
    H = Harness(False)
    elem, mt, rows, og, g = hnsw_case("vector", "l2", 1500, 12, 8, seed=31)
    ne = len(g["levels"])
    deleted = np.zeros(ne, dtype=bool)
    deleted[np.random.default_rng(3).choice(ne, 150, replace=False)] = True
    deleted[g["entry"]] = False
    image, info = P.hnsw_image(elem, 12, 8, rows[g["elem_row"]], g["levels"], g["nbr0"], g["upper_off"], g["upper"], g["entry"], deleted=deleted)
    rel, keep = H.open(image, 12, PROC[("vector", "l2")])
    gone = {P.tid_id(*P.heap_tid_of(int(e))) for e in np.nonzero(deleted)[0]}
    for qi in (3, 77, 500):
        got, _ = H.hnsw_scan(rel, elem, rows[qi], ef=60, max_items=10 ** 6)
        wi, _, _ = og.search(rows[qi], 60, ties=O.TIES_TOTAL)
        want = [P.tid_id(*P.heap_tid_of(int(e))) for e in wi if not deleted[e]]     # traversed like the reference, TIDs withheld
        assert not (set(got.tolist()) & gone)
        assert (np.array(got) == np.array(want)).mean() > 0.97 if len(got) == len(want) else False
    H.lib.h_hnsw_invalidate(rel)
    # an entry point the element pages do not hold: the scan is handed back to the CPU path, nothing raised
    image2, _ = P.hnsw_image(elem, 12, 8, rows[g["elem_row"]], g["levels"], g["nbr0"], g["upper_off"], g["upper"], g["entry"])
    b = bytearray(image2)
    import struct
    struct.pack_into("<IH", b, P.HEADER + 16, len(b) // P.BLCKSZ - 1, 200)          # entryBlkno / entryOffno: no such element
    rel2, keep2 = H.open(bytes(b), 12, PROC[("vector", "l2")])
    got, _ = H.hnsw_scan(rel2, elem, rows[0], ef=40, max_items=10)
    assert got is None and H.lib.pgstub_pinned_buffers() == 0


def _abi_build(H, elem, metric, dim, m, rows, efc, levels):
    """the C ABI called directly (mock or real): vb_hnsw_create + vb_hnsw_build + vb_hnsw_export"""
    L = H.lib
    for f in (L.vb_hnsw_create, L.vb_hnsw_build, L.vb_hnsw_export, L.vb_hnsw_free):
        f.restype = C.c_int
    L.vb_hnsw_upper_slots.restype = C.c_int64
    L.vb_hnsw_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_void_p]
    L.vb_hnsw_upper_slots.argtypes = [C.c_void_p]
    L.vb_hnsw_export.argtypes = [C.c_void_p] * 5 + [C.POINTER(C.c_int64), C.c_void_p]
    L.vb_hnsw_free.argtypes = [C.c_void_p]
    h = C.c_void_p()
    assert L.vb_hnsw_create(elem, metric, dim, m, C.byref(h)) == 0
    n = len(rows)
    lv = np.ascontiguousarray(levels, dtype=np.int32)
    assert L.vb_hnsw_build(h, np.ascontiguousarray(rows).ctypes.data_as(C.c_void_p), n, efc, 0, lv.ctypes.data_as(C.c_void_p)) == 0
    slots = L.vb_hnsw_upper_slots(h)
    out_l, nbr0 = np.empty(n, np.int32), np.empty((n, 2 * m), np.int32)
    uo, up, dup = np.empty(n, np.int64), np.full((max(slots, 1), m), -1, np.int32), np.empty(n, np.int32)
    entry = C.c_int64()
    assert L.vb_hnsw_export(h, out_l.ctypes.data_as(C.c_void_p), nbr0.ctypes.data_as(C.c_void_p), uo.ctypes.data_as(C.c_void_p),
                            up.ctypes.data_as(C.c_void_p), C.byref(entry), dup.ctypes.data_as(C.c_void_p)) == 0
    L.vb_hnsw_free(h)
    return out_l, nbr0, uo, up, entry.value, dup


@_params
@pytest.mark.parametrize("typ,metric,n,dim,m,dups", [("vector", "l2", 1500, 12, 8, 30), ("bit", "hamming", 1200, 128, 6, 0)])
def test_hnsw_build_body_materialises_the_reference_graph(real, typ, metric, n, dim, m, dups):
    """VbHnswBuildAdd / VbHnswBuildFinish: rows buffered by the build callback, graph built through the C ABI with the
    levels the glue drew, and handed to the page writer as the in-memory structure InsertTupleInMemory leaves:
    graph->head chain (newest first), neighbour arrays in stored order, duplicates as extra heap TIDs, entry point."""
    H = Harness(real)
    elem = ELEM[typ]
    mt = {"l2": O.L2_SQUARED, "hamming": O.HAMMING}[metric]
    x, _ = mixture(n, dim, 15, seed=n + m)
    rows = O.binary_quantize(O.VECTOR, x) if elem == O.BIT else x
    if dups:
        rows = np.concatenate([rows, rows[:dups]])
    nr = len(rows)
    image, _ = P.ivfflat_image(O.VECTOR, 4, np.zeros((1, 4), np.float32), np.array([0, 0]), np.zeros((0, 4), np.float32), [])
    rel, keep = H.open(image, dim, PROC[(typ, metric)])
    max_level = 12
    L = H.lib
    L.h_hnsw_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p]
    n_added, n_el, entry = C.c_int64(), C.c_int64(), C.c_int64()
    lv = np.empty(nr, np.int32)
    ht = np.full((nr, 10), -1, np.int64)
    nht = np.empty(nr, np.int32)
    nb0 = np.full((nr, 2 * m), -1, np.int64)
    upp = np.full((nr, max_level, m), -1, np.int64)
    rc = L.h_hnsw_build(rel, elem, dim, np.ascontiguousarray(rows).ctypes.data_as(C.c_void_p), nr, m, 40, 1 << 40, C.byref(n_added), C.byref(n_el),
                        lv.ctypes.data_as(C.c_void_p), ht.ctypes.data_as(C.c_void_p), nht.ctypes.data_as(C.c_void_p),
                        nb0.ctypes.data_as(C.c_void_p), upp.ctypes.data_as(C.c_void_p), max_level, C.byref(entry))
    assert rc == 0, H.err()
    assert n_added.value == nr and H.lib.pgstub_pinned_buffers() == 0
    ne = n_el.value
    tid = lambda r: P.tid_id(*P.heap_tid_of(int(r)))
    row_of = {tid(r): r for r in range(nr)}
    # the glue's level draws, per row, recovered from the elements (folded duplicates: any level, they carry no lists)
    own = np.array([row_of[int(ht[e, 0])] for e in range(ne)])
    assert np.all(np.diff(own) < 0), "graph->head chain is newest first (AddElementInMemory pushes at the head)"
    levels = np.zeros(nr, np.int32)
    levels[own] = lv[:ne]
    want_l, want_n0, want_uo, want_up, want_entry, want_dup = _abi_build(H, elem, mt, dim, m, rows, 40, levels)
    assert sorted(own.tolist()) == np.nonzero(want_dup < 0)[0].tolist()
    assert int(entry.value) == tid(want_entry)
    for e in range(ne):
        r = own[e]
        folded = sorted(np.nonzero(want_dup == r)[0].tolist())
        assert [row_of[int(t)] for t in ht[e, :nht[e]]] == [r] + folded           # own TID first, duplicates in insertion order
        got0 = [row_of[int(t)] for t in nb0[e] if t >= 0]
        assert got0 == [int(v) for v in want_n0[r] if v >= 0]
        for lc in range(1, lv[e] + 1):
            gl = [row_of[int(t)] for t in upp[e, lc - 1] if t >= 0]
            assert gl == [int(v) for v in want_up[want_uo[r] + lc - 1] if v >= 0], (e, lc)
    if dups:
        assert (want_dup >= 0).sum() >= dups // 2 and nht[:ne].max() >= 2
    # the memory budget of the in-memory phase (src/hnswbuild.c:615-624): the buffer stops accepting rows
    rc = L.h_hnsw_build(rel, elem, dim, np.ascontiguousarray(rows).ctypes.data_as(C.c_void_p), nr, m, 40, 200 * (P.row_bytes(elem, dim) + 6), C.byref(n_added),
                        C.byref(n_el), lv.ctypes.data_as(C.c_void_p), ht.ctypes.data_as(C.c_void_p), nht.ctypes.data_as(C.c_void_p),
                        nb0.ctypes.data_as(C.c_void_p), upp.ctypes.data_as(C.c_void_p), max_level, C.byref(entry))
    assert rc == 0 and n_added.value == 200 and 0 < n_el.value <= 200
