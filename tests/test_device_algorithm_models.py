"""Executable models of two device algorithms, lane for lane, checked against their specification on random inputs.  The
kernels themselves are tested on the GPU against the oracle (tests/test_gpu_hnsw.py, tests/test_gpu_ivf_one.py); these
models pin down the invariants the kernels rely on and run without a GPU:

* the bucketed visited set of the HNSW kernels (csrc/vb_hnsw.cuh, vis_insert_warp): 8-slot buckets, the lanes of one call
  arbitrated by "lowest lane wins the slot, the others take the next empty one", an id lives in the first bucket of its
  probe sequence that had room -- must behave exactly like a set, including duplicate ids inside one call and buckets
  that fill up during a call;
* the exact selection of the one-query kernels (csrc/vb_slab_select.cuh, select_exact_cta): k-th key by radix select, ties
  across the k-th place settled by the lowest indices through a bisection on the index -- must return exactly the k
  smallest (key, index) pairs.
"""
import numpy as np
from hypothesis import given, settings, strategies as st

EMPTY = 0xFFFFFFFF


def hash_u32(x):
    x &= 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def vis_insert_warp(tab, ids, want):
    """one call of the warp-collective insert: tab = flat list of slots (a power of two >= 8), ids / want = 32 lanes"""
    bmask = len(tab) // 8 - 1
    lead = {}
    for lane in range(32):                      # __match_any_sync on the id: the lowest lane carries it
        if want[lane]:
            lead.setdefault(ids[lane], lane)
    pending = [want[l] and lead[ids[l]] == l for l in range(32)]
    fresh = [False] * 32
    b = [hash_u32(ids[l]) & bmask for l in range(32)]
    while any(pending):
        snap = {}                               # every pending lane loads its bucket before any store of this pass
        for l in range(32):
            if pending[l]:
                snap[l] = list(tab[b[l] * 8:b[l] * 8 + 8])
        empt = {}
        for l, s in snap.items():
            if ids[l] in s:
                pending[l] = False
            else:
                empt[l] = [i for i in range(8) if s[i] == EMPTY]
        claim = {l: e for l, e in empt.items() if e}
        while claim:
            slot_of = {l: b[l] * 8 + e[0] for l, e in claim.items()}
            winners = {}
            for l in sorted(slot_of):           # lowest lane wins the slot
                winners.setdefault(slot_of[l], l)
            for l in list(claim):
                if winners[slot_of[l]] == l:
                    tab[slot_of[l]] = ids[l]
                    fresh[l] = True
                    pending[l] = False
                    del claim[l]
                else:
                    claim[l] = claim[l][1:]
                    if not claim[l]:
                        del claim[l]            # the bucket filled up: next bucket, next pass
        for l in range(32):
            if pending[l]:
                b[l] = (b[l] + 1) & bmask
    return fresh


def contains(tab, x):
    bmask = len(tab) // 8 - 1
    b = hash_u32(x) & bmask
    for _ in range(len(tab) // 8):
        s = tab[b * 8:b * 8 + 8]
        if x in s:
            return True
        if EMPTY in s:
            return False
        b = (b + 1) & bmask
    return False


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.sampled_from([8, 16, 64, 512]), st.integers(1, 40), st.integers(2, 5000))
def test_bucketed_visited_set_is_a_set(seed, slots, calls, id_range):
    rng = np.random.default_rng(seed)
    tab = [EMPTY] * slots
    seen = set()
    for _ in range(calls):
        if len(seen) + 32 > slots * 3 // 4:     # the kernels stop at three quarters and grow the table
            break
        ids = [int(x) for x in rng.integers(0, id_range, size=32)]
        want = [bool(x) for x in rng.integers(0, 4, size=32)]            # ~3 of 4 lanes carry an id
        fresh = vis_insert_warp(tab, ids, want)
        first = {}
        for l in range(32):
            if want[l]:
                first.setdefault(ids[l], l)
        for l in range(32):
            expect = want[l] and ids[l] not in seen and first[ids[l]] == l
            assert fresh[l] == expect, (l, ids[l])
        seen |= {ids[l] for l in range(32) if want[l]}
        stored = [x for x in tab if x != EMPTY]
        assert len(stored) == len(set(stored)) == len(seen)             # every id exactly once
        assert all(contains(tab, x) for x in seen)                      # and findable along its probe sequence


def radix_kth(keys, k):
    """the k-th smallest key (1-based) by four 8-bit passes, as ss_radix_kth"""
    prefix, mask, kk = 0, 0, k
    for p in (3, 2, 1, 0):
        shift = 8 * p
        hist = [0] * 256
        for key in keys:
            if key & mask == prefix:
                hist[(key >> shift) & 255] += 1
        cum = 0
        for b in range(256):
            if cum + hist[b] >= kk:
                break
            cum += hist[b]
        prefix |= b << shift
        mask |= 255 << shift
        kk -= cum
    return prefix


def select_exact(keys, k):
    n = len(keys)
    if n <= k:
        return sorted((keys[i], i) for i in range(n))
    tau = radix_kth(keys, k)
    less = sum(1 for x in keys if x < tau)
    eq = sum(1 for x in keys if x == tau)
    need = k - less
    assert 1 <= need <= eq
    plim = 0xFFFFFFFF
    if eq > need:
        lo, hi = 0, n - 1
        while lo < hi:
            mid = lo + (hi - lo) // 2
            if sum(1 for i in range(mid + 1) if keys[i] == tau) >= need:
                hi = mid
            else:
                lo = mid + 1
        plim = lo
    got = [(keys[i], i) for i in range(n) if keys[i] < tau or (keys[i] == tau and i <= plim)]
    assert len(got) == k
    return sorted(got)


@settings(max_examples=80, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 400), st.integers(1, 64), st.sampled_from([2, 9, 300, 2 ** 32]))
def test_exact_selection_with_ties_across_the_kth_place(seed, n, k, distinct):
    rng = np.random.default_rng(seed)
    keys = [int(x) for x in rng.integers(0, distinct, size=n, dtype=np.uint64)]
    assert select_exact(keys, k) == sorted((keys[i], i) for i in range(n))[:k]
