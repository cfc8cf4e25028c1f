"""Shared helpers for the tests: literal parsing, synthetic data, recall."""
import json
import os

import numpy as np

import oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def parse_vector(text, elem):
    """'[1,2,3]' -> float32 (vector) or uint16 half bits (halfvec); '1010' -> packed bits (MSB first)."""
    if elem == O.BIT:
        bits = np.array([c == "1" for c in text], dtype=np.uint8)
        return np.packbits(bits), len(text)   # packbits is MSB-first, zero padded tail
    vals = np.array([float(x) for x in text.strip("[]").split(",") if x.strip() != ""], dtype=np.float32)
    if elem == O.HALFVEC:
        return f32_to_half_bits(vals), vals.size
    return vals, vals.size


def f32_to_half_bits(x):
    """RNE float32 -> IEEE binary16 bit patterns (numpy's conversion is IEEE RNE;
    test_oracle_golden pins it against the oracle's and the reference's converters)."""
    with np.errstate(over="ignore"):
        return np.asarray(x, dtype=np.float32).astype(np.float16).view(np.uint16)


def half_bits_to_f32(h):
    return np.asarray(h, dtype=np.uint16).view(np.float16).astype(np.float32)


def mixture(n, dim, n_centers, seed, sigma=0.3, dtype=np.float32):
    """Gaussian mixture of SURVEY section 8(d): centres N(0,1), points centre + sigma*N(0,1)."""
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((n_centers, dim)).astype(np.float32)
    which = rng.integers(0, n_centers, size=n)
    x = centers[which] + sigma * rng.standard_normal((n, dim)).astype(np.float32)
    return x.astype(dtype), centers


def recall_at_k(got_ids, true_ids):
    """mean |got ∩ true| / k over queries."""
    hits = 0
    for g, t in zip(got_ids, true_ids):
        hits += len(set(int(x) for x in g if x >= 0) & set(int(x) for x in t))
    return hits / (len(true_ids) * len(true_ids[0]))


def build_ivf_arrays(rows, assign, lists):
    """group rows by list -> (rows_grouped, ids_grouped, offsets)"""
    order = np.argsort(assign, kind="stable")
    counts = np.bincount(assign, minlength=lists)
    offsets = np.zeros(lists + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(counts)
    return np.ascontiguousarray(rows[order]), order.astype(np.int64), offsets


def assert_same_neighbours(ids, dist, want_ids, want_dist, rtol, min_positional=0.99, boundary=2):
    """Result lists agree up to floating-point near-ties: every id the two sides share carries the same distance
    within `rtol`, ids found by one side only are confined to the cut at rank k (at most `boundary` per query),
    and the positional agreement stays above `min_positional`."""
    ids, want_ids = np.asarray(ids), np.asarray(want_ids)
    dist, want_dist = np.asarray(dist, dtype=np.float64), np.asarray(want_dist, dtype=np.float64)
    assert ids.shape == want_ids.shape
    assert (ids == want_ids).mean() > min_positional, (ids == want_ids).mean()
    for i in np.nonzero((ids != want_ids).any(axis=1))[0]:
        a = {int(t): d for t, d in zip(ids[i], dist[i]) if t >= 0}
        b = {int(t): d for t, d in zip(want_ids[i], want_dist[i]) if t >= 0}
        only = set(a) ^ set(b)
        assert len(only) <= 2 * boundary, (i, len(only))
        for t in set(a) & set(b):
            assert abs(a[t] - b[t]) <= rtol * max(abs(b[t]), 1e-30) + 1e-12, (i, t, a[t], b[t])
