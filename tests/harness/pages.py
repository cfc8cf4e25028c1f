"""Synthesise byte-exact ivfflat / hnsw index page images (test infrastructure).

Layouts: PostgreSQL's page (bufpage.h: 24-byte header, 4-byte line pointers growing up, items growing down, special
space at the end), index tuples (itup.h) and varlena headers (varatt.h, incl. the 1-byte header the server stores for
values of at most 126 bytes -- every vector of dim <= 29 in an ivfflat ENTRY tuple), and the reference's own structs:
ivfflat meta / list / entry pages (src/ivfflat.h:251-277, written by src/ivfbuild.c CreateMetaPage / CreateListPages /
the sort-ordered insert), hnsw meta page and element / neighbour tuples (src/hnsw.h:348-394, written by
src/hnswbuild.c:151-290 CreateGraphPages / WriteNeighborTuples and HnswSetNeighborTuple src/hnswutils.c:455-486).
PostgreSQL core is not in the image, so the core layouts are written from SURVEY.md Appendix B."""
import struct

import numpy as np

BLCKSZ = 8192
HEADER = 24
SPECIAL = 8
INVALID_BLOCK = 0xFFFFFFFF
IVFFLAT_MAGIC, IVFFLAT_PAGE_ID = 0x14FF1A7, 0xFF84
HNSW_MAGIC, HNSW_PAGE_ID = 0xA953A953, 0xFF90
HNSW_HEAPTIDS = 10
VB_META_VERSION_OFFSET = 64


def maxalign(x):
    return (x + 7) & ~7


class Page:
    def __init__(self, page_id):
        self.b = bytearray(BLCKSZ)
        self.lower, self.upper, self.special = HEADER, BLCKSZ - SPECIAL, BLCKSZ - SPECIAL
        self.page_id = page_id
        self.next = INVALID_BLOCK
        self.n = 0

    def free_space(self):
        space = self.upper - self.lower
        return 0 if space < 4 else space - 4

    def add_item(self, data):
        """PageAddItem: returns the 1-based offset number"""
        size = len(data)
        al = maxalign(size)
        assert self.free_space() >= al, "page full"
        self.upper -= al
        self.b[self.upper:self.upper + size] = data
        lp = (self.upper & 0x7FFF) | (1 << 15) | ((size & 0x7FFF) << 17)      # lp_off:15, lp_flags:2 = LP_NORMAL, lp_len:15
        struct.pack_into("<I", self.b, self.lower, lp)
        self.lower += 4
        self.n += 1
        return self.n

    def item_offset(self, offno):
        lp, = struct.unpack_from("<I", self.b, HEADER + 4 * (offno - 1))
        return lp & 0x7FFF

    def finish(self):
        struct.pack_into("<HHHH", self.b, 12, self.lower, self.upper, self.special, BLCKSZ | 4)
        struct.pack_into("<IHH", self.b, self.special, self.next, 0, self.page_id)
        return bytes(self.b)


def tid_bytes(block, offset):
    """ItemPointerData: bi_hi, bi_lo, ip_posid"""
    return struct.pack("<HHH", (block >> 16) & 0xFFFF, block & 0xFFFF, offset)


INVALID_TID = struct.pack("<HHH", 0xFFFF, 0xFFFF, 0)


def datum(elem, dim, payload, short_ok):
    """varlena image of a vector (0) / halfvec (1) / bit (2) value; short_ok: use the 1-byte header when it fits"""
    payload = bytes(payload)
    if elem == 2:
        body = struct.pack("<i", dim) + payload
    else:
        body = struct.pack("<hh", dim, 0) + payload
    if short_ok and len(body) + 1 <= 127:
        return bytes([((len(body) + 1) << 1) | 1]) + body
    return struct.pack("<I", (len(body) + 4) << 2) + body


def row_bytes(elem, dim):
    return dim * 4 if elem == 0 else dim * 2 if elem == 1 else (dim + 7) // 8


def heap_tid_of(i):
    """a distinct, valid heap TID per row number (block, offset >= 1)"""
    return (i // 200, i % 200 + 1)


def tid_id(block, offset):
    """the opaque int64 id the glue hands to the C ABI (VbTidToId)"""
    return (block << 16) | offset


# ------------------------------------------------------------------------------------------------ ivfflat

def ivfflat_image(elem, dim, centers, offsets, rows, row_numbers, version=0, entries_per_page=None):
    """pages of an ivfflat index holding `rows` grouped by list (offsets [lists + 1]); row_numbers[i] -> heap TID.
    Returns (bytes, info) with info['start_pages'].  entries_per_page caps the tuples per entry page (forces chains)."""
    rb = row_bytes(elem, dim)
    centers = np.ascontiguousarray(centers).view(np.uint8).reshape(len(centers), -1)
    rows = np.ascontiguousarray(rows).view(np.uint8).reshape(len(rows), -1) if len(rows) else np.zeros((0, rb), np.uint8)
    lists = len(centers)
    pages = []
    meta = Page(IVFFLAT_PAGE_ID)
    struct.pack_into("<IIHH", meta.b, HEADER, IVFFLAT_MAGIC, 1, dim, lists)
    struct.pack_into("<Q", meta.b, HEADER + VB_META_VERSION_OFFSET, version)
    meta.lower = HEADER + 12
    pages.append(meta)
    # list pages first (CreateListPages), entry pages after; list tuples are patched with their start pages below
    list_pos = []
    cur = Page(IVFFLAT_PAGE_ID)
    pages.append(cur)
    for l in range(lists):
        item = struct.pack("<II", INVALID_BLOCK, INVALID_BLOCK) + datum(elem, dim, centers[l].tobytes(), False)
        if cur.free_space() < maxalign(len(item)):
            nxt = Page(IVFFLAT_PAGE_ID)
            cur.next = len(pages)
            pages.append(nxt)
            cur = nxt
        offno = cur.add_item(item)
        list_pos.append((pages.index(cur), offno))
    start_pages = []
    for l in range(lists):
        lo, hi = int(offsets[l]), int(offsets[l + 1])
        cur = Page(IVFFLAT_PAGE_ID)
        start = len(pages)
        pages.append(cur)
        start_pages.append(start)
        for i in range(lo, hi):
            d = datum(elem, dim, rows[i].tobytes(), True)
            size = maxalign(8 + len(d))
            blk, off = heap_tid_of(int(row_numbers[i]))
            tup = tid_bytes(blk, off) + struct.pack("<H", size | 0x4000) + d     # t_info: size | INDEX_VAR_MASK
            tup += b"\0" * (size - len(tup))
            if cur.free_space() < size or (entries_per_page and cur.n >= entries_per_page):
                nxt = Page(IVFFLAT_PAGE_ID)
                cur.next = len(pages)
                pages.append(nxt)
                cur = nxt
            cur.add_item(tup)
        pg, offno = list_pos[l]
        at = pages[pg].item_offset(offno)
        struct.pack_into("<II", pages[pg].b, at, start, len(pages) - 1)         # startPage, insertPage
    return b"".join(p.finish() for p in pages), {"start_pages": start_pages, "nblocks": len(pages)}


# ------------------------------------------------------------------------------------------------ hnsw

def hnsw_image(elem, dim, m, rows, levels, nbr0, upper_off, upper, entry, heaptids=None, deleted=None, version=0,
               write_order=None):
    """pages of an hnsw index for the graph arrays (element e: rows[e], levels[e], layer-0 list nbr0[e], upper layers
    at upper[upper_off[e] + lc - 1]).  heaptids[e] = list of row numbers carried by the element (default [e]).
    write_order: element numbers in the order CreateGraphPages writes them (the reference: newest first)."""
    n = len(rows)
    rows = np.ascontiguousarray(rows).view(np.uint8).reshape(n, -1)
    order = list(write_order) if write_order is not None else list(range(n - 1, -1, -1))
    pages = []
    meta = Page(HNSW_PAGE_ID)
    pages.append(meta)
    cur = Page(HNSW_PAGE_ID)
    pages.append(cur)
    where = {}          # element -> (blk, off, nblk, noff)
    slots = {}          # element -> (page index, item offset in page, size) of the neighbour tuple
    max_size = BLCKSZ - maxalign(HEADER) - maxalign(SPECIAL) - 4

    def new_page():
        nonlocal cur
        nxt = Page(HNSW_PAGE_ID)
        cur.next = len(pages)
        pages.append(nxt)
        cur = nxt

    for e in order:
        d = datum(elem, dim, rows[e].tobytes(), False)
        etup_size = maxalign(4 + 6 * HNSW_HEAPTIDS + 6 + 2 + len(d))      # offsetof(HnswElementTupleData, data) = 72
        ntup_size = maxalign(4 + 6 * (int(levels[e]) + 2) * m)
        combined = etup_size + ntup_size + 4
        if cur.free_space() < etup_size or (combined <= max_size and cur.free_space() < combined):
            new_page()
        blk, off = len(pages) - 1, cur.n + 1
        if combined <= max_size:
            nblk, noff = blk, off + 1
        else:
            nblk, noff = blk + 1, 1
        hts = [e] if heaptids is None else list(heaptids[e])
        is_del = bool(deleted is not None and deleted[e])
        tids = b"".join(tid_bytes(*heap_tid_of(int(h))) for h in hts) + INVALID_TID * (HNSW_HEAPTIDS - len(hts))
        etup = struct.pack("<BBBB", 1, int(levels[e]), 1 if is_del else 0, 1) + tids + tid_bytes(nblk, noff) + struct.pack("<H", 0) + d
        etup += b"\0" * (etup_size - len(etup))
        assert cur.add_item(etup) == off
        if cur.free_space() < ntup_size:
            new_page()
        assert (len(pages) - 1, cur.n + 1) == (nblk, noff), ((len(pages) - 1, cur.n + 1), (nblk, noff))
        cur.add_item(b"\0" * ntup_size)
        slots[e] = (len(pages) - 1, cur.item_offset(cur.n), ntup_size)
        where[e] = (blk, off, nblk, noff)
    # WriteNeighborTuples: HnswSetNeighborTuple order -- layer `level` first, layer 0 last, m (2m at layer 0) slots each
    for e in order:
        lv = int(levels[e])
        out = bytearray()
        count = 0
        for lc in range(lv, -1, -1):
            lm = 2 * m if lc == 0 else m
            lst = nbr0[e] if lc == 0 else upper[int(upper_off[e]) + lc - 1]
            for i in range(lm):
                v = int(lst[i]) if i < len(lst) else -1
                out += tid_bytes(where[v][0], where[v][1]) if v >= 0 else INVALID_TID
                count += 1
        ntup = struct.pack("<BBH", 2, 1, count) + bytes(out)
        pg, at, size = slots[e]
        pages[pg].b[at:at + len(ntup)] = ntup
    eb, eo = (where[entry][0], where[entry][1]) if entry is not None and entry >= 0 else (INVALID_BLOCK, 0)
    struct.pack_into("<IIIHHIHhI", meta.b, HEADER, HNSW_MAGIC, 1, dim, m, 64, eb, eo, int(levels[entry]) if entry is not None and entry >= 0 else -1,
                     len(pages) - 1)
    struct.pack_into("<Q", meta.b, HEADER + VB_META_VERSION_OFFSET, version)
    meta.lower = HEADER + 28
    return b"".join(p.finish() for p in pages), {"where": where, "nblocks": len(pages)}
