"""pgvector_b200 -- host-side mirror of pgvector's distance / index-scan interface
over libvecb200.so (hand-written sm_100a CUDA behind the C ABI of include/vecb200.h).

PostgreSQL is not available in this image, so the reference's C host code
(index AM callbacks) cannot be linked here; the extension-side glue is under
``pgvector_b200/ext`` (written against the PostgreSQL API) and this module is
the thin Python mirror of the same operator / opclass surface that the parity
tests and ``bench.py`` drive.  Names follow the reference: operators
(``l2_distance`` ... ``jaccard_distance``, src/vector.c:576-750,
src/bitvec.c:45-70), opclasses (``vector_l2_ops`` ..., sql/vector.sql:406-446,
819-911), ``IvfflatIndex`` (src/ivfscan.c) and ``HnswIndex`` (src/hnswscan.c).

Everything computes on the GPU through the C ABI; nothing here falls back to
numpy / torch math, and nothing imports the test oracle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import sparsevec  # noqa: F401  (sparsevec functions, CSR tables and the exact scan over them)
from ._lib import VecB200Error, load  # noqa: F401
from .sparsevec import SparseRows, SparseTable, SparseVector  # noqa: F401

VECTOR, HALFVEC, BIT = 0, 1, 2
L2_SQUARED, NEG_IP, COSINE, L1, HAMMING, JACCARD, L2, IP, SPHERICAL = range(9)

_NP = {VECTOR: np.float32, HALFVEC: np.uint16, BIT: np.uint8}
_TYPE_NAME = {VECTOR: "vector", HALFVEC: "halfvec", BIT: "bit"}

# opclass -> (element type, proc-1 metric the index evaluates, normalise rows/query?, k-means metric)
# (sql/vector.sql:406-446, 819-866, 894-911)
OPCLASSES = {
    "vector_l2_ops": (VECTOR, L2_SQUARED, False, L2),
    "vector_ip_ops": (VECTOR, NEG_IP, False, SPHERICAL),
    "vector_cosine_ops": (VECTOR, NEG_IP, True, SPHERICAL),
    "vector_l1_ops": (VECTOR, L1, False, None),
    "halfvec_l2_ops": (HALFVEC, L2_SQUARED, False, L2),
    "halfvec_ip_ops": (HALFVEC, NEG_IP, False, SPHERICAL),
    "halfvec_cosine_ops": (HALFVEC, NEG_IP, True, SPHERICAL),
    "halfvec_l1_ops": (HALFVEC, L1, False, None),
    "bit_hamming_ops": (BIT, HAMMING, False, HAMMING),
    "bit_jaccard_ops": (BIT, JACCARD, False, None),
}


def init(device: int = 0):
    _lib.check(load().vb_init(device))


def stream_handle() -> int:
    """cudaStream_t of the library (int) -- wrap with torch.cuda.ExternalStream for event timing."""
    return int(load().vb_stream() or 0)


def launch_count() -> int:
    return int(load().vb_launch_count())


PROF_SCAN_ITEMS, PROF_SCAN_LISTS, PROF_TOPK, PROF_ASSIGN, PROF_HNSW, PROF_LIST_TC, PROF_CENTRE_TC = range(7)


def prof_enable(on=True):
    _lib.check(load().vb_prof_enable(1 if on else 0))


def prof_read(kernel):
    """(total milliseconds, launches) of the bracketed kernel class since the last read."""
    ms, n = C.c_double(), C.c_int64()
    _lib.check(load().vb_prof_read(kernel, C.byref(ms), C.byref(n)))
    return ms.value, n.value


def synchronize():
    _lib.check(load().vb_synchronize())


def _after_torch(*tensors):
    """The library launches on its own non-blocking stream: before a call that reads torch CUDA tensors, make that
    stream wait for whatever torch's current stream has enqueued so far (the tensors' producers)."""
    if any(_is_torch(t) and t.is_cuda for t in tensors):
        import torch
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        _lib.check(load().vb_stream_wait_event(C.c_void_p(ev.cuda_event)))


def _host(elem, a):
    """host array in the payload layout of the type.  halfvec rows are IEEE binary16 BIT PATTERNS (uint16): float16
    arrays are reinterpreted, float32/64 arrays are rounded to half (RNE, like the vector -> halfvec cast,
    src/halfvec.c:540-555, without its overflow check) -- never value-cast to integers."""
    a = np.asarray(a)
    if elem == HALFVEC and a.dtype != np.uint16:
        if a.dtype == np.float16:
            a = a.view(np.uint16)
        elif a.dtype in (np.float32, np.float64):
            with np.errstate(over="ignore"):
                a = a.astype(np.float16).view(np.uint16)
        else:
            raise TypeError(f"halfvec rows must be uint16 bit patterns or a float array, not {a.dtype}")
    return np.ascontiguousarray(a, dtype=_NP[elem])


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(a.data_ptr())  # torch tensor


def _is_torch(a):
    return a is not None and not isinstance(a, np.ndarray) and hasattr(a, "data_ptr")


def _dim_of(elem, a, dim):
    if dim is not None:
        return int(dim)
    return int(a.shape[-1]) * (8 if elem == BIT else 1)


def _check_dims(elem, da, db):
    # CheckDims (src/vector.c:70-77, src/halfvec.c:74-81, src/bitvec.c:33-40)
    if da != db:
        kind = {VECTOR: "vector dimensions", HALFVEC: "halfvec dimensions", BIT: "bit lengths"}[elem]
        raise ValueError(f"different {kind} {da} and {db}")


# --------------------------------------------------------------------- operators

def distance_batch(elem, metric, q, rows, dim=None, q_dim=None):
    """float8 distances of one query against n rows (host arrays), as the fmgr wrapper returns them."""
    rows = _host(elem, rows)
    if rows.ndim == 1:
        rows = rows.reshape(1, -1)
    d = _dim_of(elem, rows, dim)
    if q is not None:
        q = _host(elem, q)
        _check_dims(elem, q_dim if q_dim is not None else _dim_of(elem, q, dim), d)
    out = np.empty(rows.shape[0], dtype=np.float64)
    _lib.check(load().vb_distance_batch(elem, metric, d, _ptr(q), _ptr(rows), rows.shape[0], _ptr(out)))
    return out


def l2_distance(a, rows, elem=VECTOR, **kw):
    return distance_batch(elem, L2, a, rows, **kw)


def l2_squared_distance(a, rows, elem=VECTOR, **kw):
    return distance_batch(elem, L2_SQUARED, a, rows, **kw)


def inner_product(a, rows, elem=VECTOR, **kw):
    return distance_batch(elem, IP, a, rows, **kw)


def negative_inner_product(a, rows, elem=VECTOR, **kw):
    return distance_batch(elem, NEG_IP, a, rows, **kw)


def cosine_distance(a, rows, elem=VECTOR, **kw):
    return distance_batch(elem, COSINE, a, rows, **kw)


def l1_distance(a, rows, elem=VECTOR, **kw):
    return distance_batch(elem, L1, a, rows, **kw)


def spherical_distance(a, rows, elem=VECTOR, **kw):
    return distance_batch(elem, SPHERICAL, a, rows, **kw)


def hamming_distance(a, rows, dim=None, **kw):
    return distance_batch(BIT, HAMMING, a, rows, dim=dim, **kw)


def jaccard_distance(a, rows, dim=None, **kw):
    return distance_batch(BIT, JACCARD, a, rows, dim=dim, **kw)


# --------------------------------------------------------------------- resident table / exact scan

class Table:
    """[n x dim] rows resident in HBM."""

    def __init__(self, elem, dim):
        self.elem, self.dim = elem, int(dim)
        h = C.c_void_p()
        _lib.check(load().vb_table_create(elem, self.dim, C.byref(h)))
        self.h = h

    def append(self, rows):
        if _is_torch(rows):
            _after_torch(rows)
            _lib.check(load().vb_table_append_dev(self.h, _ptr(rows), rows.shape[0]))
        else:
            rows = _host(self.elem, rows)
            _lib.check(load().vb_table_append(self.h, _ptr(rows), rows.shape[0]))
        return self

    def __len__(self):
        return int(load().vb_table_rows(self.h))

    def device_rows(self):
        """(device pointer of row 0, padded row stride in bytes): a read-only view of the resident rows"""
        stride = C.c_size_t()
        p = load().vb_table_device_rows(self.h, C.byref(stride))
        return int(p or 0), int(stride.value)

    def exact_topk(self, metric, queries, k):
        """ORDER BY v <op> q LIMIT k without an index."""
        if _is_torch(queries):
            import torch
            nq = queries.shape[0]
            ids = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
            dist = torch.empty((nq, k), dtype=torch.float32, device=queries.device)
            _after_torch(queries)
            _lib.check(load().vb_exact_topk_dev(self.h, metric, _ptr(queries), nq, k, _ptr(ids), _ptr(dist)))
            synchronize()   # the library runs on its own stream; results are handed back complete
            return ids, dist
        queries = _host(self.elem, queries)
        if queries.ndim == 1:
            queries = queries.reshape(1, -1)
        nq = queries.shape[0]
        ids = np.empty((nq, k), dtype=np.int64)
        dist = np.empty((nq, k), dtype=np.float64)
        _lib.check(load().vb_exact_topk(self.h, metric, _ptr(queries), nq, k, _ptr(ids), _ptr(dist)))
        return ids, dist

    def exact_topk_sharded(self, metric, queries_dev, k, id_offset):
        """exact top-k over a row-sharded table (collective over the library's communicator); torch CUDA tensors"""
        import torch
        nq = queries_dev.shape[0]
        ids = torch.empty((nq, k), dtype=torch.int64, device=queries_dev.device)
        dist = torch.empty((nq, k), dtype=torch.float32, device=queries_dev.device)
        _after_torch(queries_dev)
        _lib.check(load().vb_exact_topk_sharded_dev(self.h, metric, _ptr(queries_dev), nq, k, int(id_offset), _ptr(ids), _ptr(dist)))
        synchronize()
        return ids, dist

    def free(self):
        if self.h:
            load().vb_table_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# --------------------------------------------------------------------- IVFFlat

class IvfflatIndex:
    """Device image of an ivfflat index and its scan (src/ivfscan.c).

    ``probes`` mirrors the ivfflat.probes GUC (src/ivfflat.c:45-47).  For the cosine opclasses (``self.normalize``)
    the reference stores l2_normalize'd rows (src/ivfbuild.c:174-180) and normalises the query once per scan
    (src/ivfscan.c:222-229); this image takes rows and centres as stored, and ``prepare_query`` applies the
    query-side normalisation."""

    def __init__(self, opclass, dim, lists):
        self.opclass = opclass
        self.elem, self.metric, self.normalize, self.kmeans_metric = OPCLASSES[opclass]
        if self.metric not in (L2_SQUARED, NEG_IP, HAMMING):
            raise ValueError(f"operator class {opclass} is not supported by ivfflat")
        self.dim, self.lists = int(dim), int(lists)
        self.probes = 1  # IVFFLAT_DEFAULT_PROBES
        h = C.c_void_p()
        _lib.check(load().vb_ivf_create(self.elem, self.metric, self.dim, self.lists, C.byref(h)))
        self.h = h

    def load(self, centers, list_offsets, rows, ids=None):
        off = np.ascontiguousarray(list_offsets, dtype=np.int64)
        assert off.shape[0] == self.lists + 1
        self._off = off
        if _is_torch(rows):
            self._keep = (centers, rows, ids)
            _after_torch(centers, rows, ids)
            _lib.check(load().vb_ivf_load_dev(self.h, _ptr(centers), _ptr(off), _ptr(rows), _ptr(ids)))
        else:
            centers = _host(self.elem, centers)
            rows = _host(self.elem, rows)
            ids = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
            _lib.check(load().vb_ivf_load(self.h, _ptr(centers), _ptr(off), _ptr(rows), _ptr(ids)))
        return self

    def prepare_query(self, q):
        """what ivfflatgettuple does to the ORDER BY value before the scan (src/ivfscan.c:213-231): l2_normalize for
        the cosine opclasses, identity otherwise."""
        return l2_normalize(q, self.elem) if self.normalize else q

    def load_by_list(self, centers, lists):
        """the same image loaded one list at a time: `lists` = iterable of (list number, rows, ids), ascending"""
        centers = _host(self.elem, centers)
        _lib.check(load().vb_ivf_begin_load(self.h, _ptr(centers)))
        off = np.zeros(self.lists + 1, dtype=np.int64)
        for l, rows, ids in lists:
            rows = _host(self.elem, rows)
            ids = np.ascontiguousarray(ids, dtype=np.int64)
            _lib.check(load().vb_ivf_load_list(self.h, int(l), _ptr(rows), _ptr(ids), rows.shape[0]))
            off[l + 1] = rows.shape[0]
        _lib.check(load().vb_ivf_end_load(self.h))
        self._off = np.concatenate([[0], np.cumsum(off[1:])]).astype(np.int64)
        return self

    def replace_list(self, l, rows, ids):
        """swap one list of the loaded image (what an insert into / a vacuum of that list needs)"""
        rows = _host(self.elem, rows)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        _lib.check(load().vb_ivf_replace_list(self.h, int(l), _ptr(rows), _ptr(ids), rows.shape[0]))
        delta = rows.shape[0] - int(self._off[l + 1] - self._off[l])
        self._off = self._off.copy()
        self._off[l + 1:] += delta
        return self

    def scan_lists(self, queries, max_probes=None):
        """GetScanLists: nearest lists per query, ascending."""
        mp = int(max_probes or self.probes)
        if queries is None:
            nq, q = 1, None
        else:
            q = _host(self.elem, queries)
            if q.ndim == 1:
                q = q.reshape(1, -1)
            nq = q.shape[0]
        lists = np.empty((nq, mp), dtype=np.int32)
        dist = np.empty((nq, mp), dtype=np.float64)
        _lib.check(load().vb_ivf_scan_lists(self.h, _ptr(q), nq, mp, _ptr(lists), _ptr(dist)))
        return lists, dist

    def scan_items(self, q, lists, cap=None):
        """GetScanItems for one query: every row of `lists`, sorted by distance."""
        lists = np.ascontiguousarray(lists, dtype=np.int32)
        total = int(sum(self._off[l + 1] - self._off[l] for l in lists))
        cap = total if cap is None else min(int(cap), total)
        ids = np.empty(max(cap, 1), dtype=np.int64)
        dist = np.empty(max(cap, 1), dtype=np.float64)
        n = C.c_int64()
        qh = None if q is None else _host(self.elem, q)
        _lib.check(load().vb_ivf_scan_items(self.h, _ptr(qh), _ptr(lists), len(lists), cap, _ptr(ids), _ptr(dist), C.byref(n)))
        return ids[:cap], dist[:cap], int(n.value)

    def search(self, queries, k, probes=None):
        """first batch of ivfflatgettuple for many queries: k nearest of the probed lists."""
        p = int(probes or self.probes)
        if _is_torch(queries):
            import torch
            nq = queries.shape[0]
            ids = torch.empty((nq, k), dtype=torch.int64, device=queries.device)
            dist = torch.empty((nq, k), dtype=torch.float32, device=queries.device)
            _after_torch(queries)
            _lib.check(load().vb_ivf_search_dev(self.h, _ptr(queries), nq, p, k, _ptr(ids), _ptr(dist)))
            synchronize()   # (search_into is the asynchronous variant)
            return ids, dist
        q = _host(self.elem, queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.int64)
        dist = np.empty((nq, k), dtype=np.float64)
        _lib.check(load().vb_ivf_search(self.h, _ptr(q), nq, p, k, _ptr(ids), _ptr(dist)))
        return ids, dist

    def search_into(self, queries_dev, k, probes, ids_dev, dist_dev):
        """asynchronous device-resident search into preallocated torch tensors (bench inner loop): enqueued on the
        library stream after torch's current stream; the caller synchronises (pv.synchronize()) before reading."""
        _after_torch(queries_dev)
        _lib.check(load().vb_ivf_search_dev(self.h, _ptr(queries_dev), queries_dev.shape[0], int(probes), int(k),
                                            _ptr(ids_dev), _ptr(dist_dev)))

    def search_sharded_into(self, queries_dev, k, probes, ids_dev, dist_dev):
        """list-sharded search over the library's communicator (collective; every rank gets the full result)"""
        _after_torch(queries_dev)
        _lib.check(load().vb_ivf_search_sharded_dev(self.h, _ptr(queries_dev), queries_dev.shape[0], int(probes), int(k),
                                                    _ptr(ids_dev), _ptr(dist_dev)))

    def search_sharded_host_into(self, queries, k, probes, ids, dist):
        """list-sharded search, host buffers in and out (int64 ids, float64 distances)"""
        _lib.check(load().vb_ivf_search_sharded(self.h, _ptr(queries), queries.shape[0], int(probes), int(k), _ptr(ids), _ptr(dist)))

    def search_host_into(self, queries, k, probes, ids, dist):
        _lib.check(load().vb_ivf_search(self.h, _ptr(queries), queries.shape[0], int(probes), int(k), _ptr(ids), _ptr(dist)))

    def prefetch_queries(self, queries, slot):
        """start the H2D copy of the next batch (pinned host array) into slot 0 / 1; returns at once"""
        _lib.check(load().vb_ivf_prefetch_queries(self.h, _ptr(queries), queries.shape[0], int(slot)))

    def search_prefetched_into(self, slot, k, probes, ids, dist):
        """search the batch prefetched into `slot`; host outputs (int64 ids, float64 distances)"""
        _lib.check(load().vb_ivf_search_prefetched(self.h, int(slot), int(probes), int(k), _ptr(ids), _ptr(dist)))

    def last_scan_bytes(self):
        return int(load().vb_ivf_last_scan_bytes(self.h))

    def last_candidates(self):
        return int(load().vb_ivf_last_candidates(self.h))

    def tc_level1_fallbacks(self):
        """queries the hi-plane-only filter level could not certify (their batches were repeated with both planes)"""
        return int(load().vb_ivf_tc_level1_fallbacks(self.h))

    def tc_fallbacks(self):
        """queries re-run exactly because the tensor-core filter could not certify them (scan_impl = 4)"""
        return int(load().vb_ivf_tc_fallbacks(self.h))

    def free(self):
        if self.h:
            load().vb_ivf_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# --------------------------------------------------------------------- IVFFlat build

def make_allreduce(fn):
    """wrap a python callable (ptr, count, dtype_code) -> None as the C hook."""
    def _cb(buf, count, dtype, _ctx):
        try:
            fn(buf, count, dtype)
            return 0
        except Exception:  # pragma: no cover - surfaced as VB error
            return -1
    return _lib.ALLREDUCE_FN(_cb)


def kmeans(samples: Table, kmeans_metric, init_centers, max_iter=500, seed=42, allreduce=None):
    """IvfflatKmeans (src/ivfkmeans.c:553-570) from given initial centres."""
    centers = _host(samples.elem, init_centers).copy()
    k = centers.shape[0]
    iters = C.c_int()
    cb = make_allreduce(allreduce) if allreduce is not None else None
    _lib.check(load().vb_kmeans(samples.h, kmeans_metric, _ptr(centers), k, max_iter, seed,
                                C.cast(cb, C.c_void_p) if cb is not None else None, None, C.byref(iters)))
    return centers, iters.value


def kmeans_pp_init(samples: Table, kmeans_metric, k, seed=42):
    raw = (samples.dim + 7) // 8 if samples.elem == BIT else samples.dim
    centers = np.empty((k, raw), dtype=_NP[samples.elem])
    _lib.check(load().vb_kmeans_pp_init(samples.h, kmeans_metric, _ptr(centers), k, seed))
    return centers


def kmeans_pp_stats():
    """(skipped by the triangle rule, stopped by the bf16 bound, re-scored exactly) of the last k-means++ seeding"""
    out = np.zeros(3, dtype=np.int64)
    _lib.check(load().vb_kmeans_pp_stats(_ptr(out)))
    return tuple(int(x) for x in out)


def kmeans_pp_init_draws(samples: Table, kmeans_metric, k, first_row, u):
    """InitCenters (src/ivfkmeans.c:23-91) with the caller's draws; returns (centres, picked sample rows)."""
    raw = (samples.dim + 7) // 8 if samples.elem == BIT else samples.dim
    centers = np.empty((k, raw), dtype=_NP[samples.elem])
    u = np.ascontiguousarray(u, dtype=np.float64)
    picked = np.empty(k, dtype=np.int64)
    _lib.check(load().vb_kmeans_pp_init_draws(samples.h, kmeans_metric, _ptr(centers), k, int(first_row), _ptr(u), _ptr(picked)))
    return centers, picked


def assign(rows: Table, metric, centers):
    """AddTupleToSort's nearest-centre pass (src/ivfbuild.c:161-219)."""
    if _is_torch(centers):
        import torch
        out = torch.empty(len(rows), dtype=torch.int32, device=centers.device)
        _after_torch(centers)
        _lib.check(load().vb_assign_dev(rows.h, metric, _ptr(centers), centers.shape[0], _ptr(out)))
        return out
    centers = _host(rows.elem, centers)
    out = np.empty(len(rows), dtype=np.int32)
    _lib.check(load().vb_assign(rows.h, metric, _ptr(centers), centers.shape[0], _ptr(out)))
    return out


# --------------------------------------------------------------------- HNSW

class HnswIndex:
    """Device image of an hnsw index and its scan (src/hnswscan.c, src/hnswutils.c:824-987).

    ``ef_search`` mirrors the hnsw.ef_search GUC (src/hnsw.c:93-95)."""

    def __init__(self, opclass, dim, m=16):
        self.opclass = opclass
        self.elem, self.metric, self.normalize, _ = OPCLASSES[opclass]
        self.dim, self.m = int(dim), int(m)
        self.ef_search = 40  # HNSW_DEFAULT_EF_SEARCH
        h = C.c_void_p()
        _lib.check(load().vb_hnsw_create(self.elem, self.metric, self.dim, self.m, C.byref(h)))
        self.h = h

    def load(self, rows, levels, nbr0, upper_off, upper, entry):
        rows = _host(self.elem, rows)
        levels = np.ascontiguousarray(levels, dtype=np.int32)
        nbr0 = np.ascontiguousarray(nbr0, dtype=np.int32)
        upper_off = np.ascontiguousarray(upper_off, dtype=np.int64)
        upper = np.ascontiguousarray(upper, dtype=np.int32)
        slots = upper.shape[0] if upper.size else 0
        self.n = rows.shape[0]
        _lib.check(load().vb_hnsw_load(self.h, _ptr(rows), rows.shape[0], _ptr(levels), _ptr(nbr0), _ptr(upper_off),
                                       _ptr(upper) if slots else None, slots, int(entry)))
        return self

    def build(self, rows, ef_construction=64, seed=42, levels=None):
        """CREATE INDEX on the device (src/hnswbuild.c:437-480 in batches): row i becomes element i."""
        lv = None if levels is None else np.ascontiguousarray(levels, dtype=np.int32)
        if _is_torch(rows):
            _after_torch(rows)
            self.n = rows.shape[0]
            _lib.check(load().vb_hnsw_build_dev(self.h, _ptr(rows), self.n, int(ef_construction), int(seed), _ptr(lv)))
        else:
            rows = _host(self.elem, rows)
            self.n = rows.shape[0]
            _lib.check(load().vb_hnsw_build(self.h, _ptr(rows), self.n, int(ef_construction), int(seed), _ptr(lv)))
        return self

    def export(self):
        """the graph as arrays (the layout load() takes, plus dup_of): what the page writer consumes"""
        n, m = int(load().vb_hnsw_rows(self.h)), self.m
        slots = int(load().vb_hnsw_upper_slots(self.h))
        levels = np.empty(n, dtype=np.int32)
        nbr0 = np.empty((n, 2 * m), dtype=np.int32)
        upper_off = np.empty(n, dtype=np.int64)
        upper = np.full((max(slots, 1), m), -1, dtype=np.int32)
        dup_of = np.empty(n, dtype=np.int32)
        entry = C.c_int64(-1)
        _lib.check(load().vb_hnsw_export(self.h, _ptr(levels), _ptr(nbr0), _ptr(upper_off), _ptr(upper), C.byref(entry), _ptr(dup_of)))
        e = int(entry.value)
        return dict(levels=levels, nbr0=nbr0, upper_off=upper_off, upper=upper[:slots], entry=e,
                    entry_level=int(levels[e]) if e >= 0 else -1, m=m, dup_of=dup_of)

    def search(self, queries, k=None, ef_search=None):
        ef = int(ef_search or self.ef_search)
        k = int(k or ef)
        q = _host(self.elem, queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.int64)
        dist = np.empty((nq, k), dtype=np.float64)
        nd = np.empty(nq, dtype=np.int64)
        _lib.check(load().vb_hnsw_search(self.h, _ptr(q), nq, ef, k, _ptr(ids), _ptr(dist), _ptr(nd)))
        return ids, dist, nd

    def search_into(self, queries_dev, k, ef, ids_dev, dist_dev, nd_dev=None):
        _after_torch(queries_dev)
        _lib.check(load().vb_hnsw_search_dev(self.h, _ptr(queries_dev), queries_dev.shape[0], int(ef), int(k),
                                             _ptr(ids_dev), _ptr(dist_dev), _ptr(nd_dev)))

    def iterative_scan(self, queries, ef_search=None, max_scan_tuples=20000):
        """hnsw.iterative_scan for a batch of queries: an HnswScan whose next_batch() mirrors ResumeScanItems
        (src/hnswscan.c:62-87); max_scan_tuples mirrors hnsw.max_scan_tuples (src/hnsw.c:101-105)."""
        return HnswScan(self, queries, int(ef_search or self.ef_search), int(max_scan_tuples))

    def free(self):
        if self.h:
            load().vb_hnsw_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class HnswScan:
    """One iterative index scan per query (src/hnswscan.c:228-340): next_batch() returns (ids, distances, counts) of
    the next <= ef_search elements of every query, nearest first; counts == 0 marks an exhausted scan."""

    def __init__(self, index, queries, ef, max_scan_tuples):
        q = _host(index.elem, queries)
        if q.ndim == 1:
            q = q.reshape(1, -1)
        self.index, self.nq, self.ef = index, q.shape[0], ef
        h = C.c_void_p()
        _lib.check(load().vb_hnsw_scan_begin(index.h, _ptr(q), self.nq, ef, max_scan_tuples, C.byref(h)))
        self.h = h

    def next_batch(self):
        ids = np.empty((self.nq, self.ef), dtype=np.int64)
        dist = np.empty((self.nq, self.ef), dtype=np.float64)
        cnt = np.empty(self.nq, dtype=np.int32)
        _lib.check(load().vb_hnsw_scan_next(self.h, _ptr(ids), _ptr(dist), _ptr(cnt)))
        return ids, dist, cnt

    def tuples(self):
        t = np.empty(self.nq, dtype=np.int64)
        _lib.check(load().vb_hnsw_scan_tuples(self.h, _ptr(t)))
        return t

    def tuples_of(self, query=0, strict=False, limit=None):
        """what hnswgettuple hands the executor for one query, in order: (element, distance) pairs; strict mirrors
        hnsw.iterative_scan = strict_order (elements nearer than one already returned are skipped, :316-322)"""
        out = []
        prev = -np.inf
        while limit is None or len(out) < limit:
            ids, dist, cnt = self.next_batch()
            if cnt[query] == 0:
                break
            for j in range(int(cnt[query])):
                d = float(dist[query, j])
                if strict:
                    if d < prev:
                        continue
                    prev = d
                out.append((int(ids[query, j]), d))
        return out if limit is None else out[:limit]

    def close(self):
        if self.h:
            load().vb_hnsw_scan_end(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# --------------------------------------------------------------------- communicator (one process per GPU)

def comm_unique_id() -> bytes:
    """the 128-byte NCCL id one rank creates and the host hands to the others"""
    buf = C.create_string_buffer(128)
    _lib.check(load().vb_comm_unique_id(buf, 128))
    return buf.raw


def comm_init(id_bytes: bytes, rank: int, world: int):
    """collective: create the library's own NCCL communicator on its device"""
    _lib.check(load().vb_comm_init(C.c_char_p(id_bytes), int(rank), int(world)))


def comm_free():
    _lib.check(load().vb_comm_free())


def comm_world() -> int:
    return int(load().vb_comm_world())


def tc_traffic(on=True, read=False):
    """traffic accounting of the tensor-core filter launches; read=True returns and resets the 8 counters"""
    out = np.zeros(8, dtype=np.int64) if read else None
    _lib.check(load().vb_ivf_tc_traffic(1 if on else 0, _ptr(out)))
    return out


def set_tensor_cores(on: bool):
    """False forces the exact fp32 CUDA-core assign kernel (parity tests); True (default) uses tcgen05."""
    _lib.check(load().vb_set_tensor_cores(1 if on else 0))


def last_assign_rechecked() -> int:
    return int(load().vb_last_assign_rechecked())


def set_option(name: str, value: int):
    """tuning switches of the library: "scan_impl" (0 = LDG kernel, 1 = bulk-copy/TMA kernel), "tensor_cores"."""
    _lib.check(load().vb_set_option(name.encode(), int(value)))


# --------------------------------------------------------------------- row transforms

def vector_norm(rows, elem=VECTOR):
    """vector_norm / l2_norm of every row (src/vector.c:767-780, src/halfvec.c:703-720)."""
    rows = _host(elem, rows)
    single = rows.ndim == 1
    r2 = rows.reshape(1, -1) if single else rows
    out = np.empty(r2.shape[0], dtype=np.float64)
    _lib.check(load().vb_norm_batch(elem, r2.shape[1], _ptr(r2), r2.shape[0], _ptr(out)))
    return out[0] if single else out


def l2_normalize(rows, elem=VECTOR):
    """l2_normalize of every row (src/vector.c:785-819, src/halfvec.c:725-759); raises OverflowError like the reference."""
    rows = _host(elem, rows)
    single = rows.ndim == 1
    r2 = np.ascontiguousarray(rows.reshape(1, -1) if single else rows)
    out = np.empty_like(r2)
    try:
        _lib.check(load().vb_l2_normalize_batch(elem, r2.shape[1], _ptr(r2), r2.shape[0], _ptr(out)))
    except VecB200Error as e:
        if "overflow" in str(e):
            raise OverflowError("value out of range: overflow") from None
        raise
    return out[0] if single else out


def vector_to_halfvec(rows):
    """vector::halfvec (src/halfvec.c:540-555): RNE to half bit patterns; raises like the reference on overflow"""
    rows = _host(VECTOR, rows)
    single = rows.ndim == 1
    r2 = np.ascontiguousarray(rows.reshape(1, -1) if single else rows)
    out = np.empty(r2.shape, dtype=np.uint16)
    try:
        _lib.check(load().vb_vector_to_halfvec_batch(r2.shape[1], _ptr(r2), r2.shape[0], _ptr(out)))
    except VecB200Error as e:
        if "out of range for type halfvec" in str(e):
            raise ValueError(str(e).split(": ", 1)[1]) from None
        raise
    return out[0] if single else out


def halfvec_to_vector(rows):
    """halfvec::vector: exact widening"""
    rows = _host(HALFVEC, rows)
    single = rows.ndim == 1
    r2 = np.ascontiguousarray(rows.reshape(1, -1) if single else rows)
    out = np.empty(r2.shape, dtype=np.float32)
    _lib.check(load().vb_halfvec_to_vector_batch(r2.shape[1], _ptr(r2), r2.shape[0], _ptr(out)))
    return out[0] if single else out


def binary_quantize(rows, elem=VECTOR):
    """binary_quantize of every row (src/vector.c:952-978): packed bits, MSB first."""
    rows = _host(elem, rows)
    single = rows.ndim == 1
    r2 = np.ascontiguousarray(rows.reshape(1, -1) if single else rows)
    out = np.empty((r2.shape[0], (r2.shape[1] + 7) // 8), dtype=np.uint8)
    _lib.check(load().vb_binary_quantize_batch(elem, r2.shape[1], _ptr(r2), r2.shape[0], _ptr(out)))
    return out[0] if single else out
