"""
CPU check of the bookkeeping of the rank-tracking moving median (gordo_b200/csrc/smooth.cu, smm_rank_kernel): the
same state machine restated in Python -- unsorted ring of order-preserving integer keys, one pass per output that
yields pred / second pred / succ of the previous rank-k key r and #{key == r}, the carried count below / above r --
against pandas `rolling(w).median()` (diff.py:302-308 is what the reference calls) on ties, signed zeros, missing
values and +-inf.  The GPU tests (tests/test_gpu_smooth.py) check the kernel itself; this one pins the logic the
kernel implements, on a box without a GPU.
"""
import numpy as np
import pandas as pd
import pytest

MISSING = 0xFFFFFFFF


def f2key(x):
    u = int(np.float32(x).view(np.uint32))
    return (~u) & 0xFFFFFFFF if u & 0x80000000 else u | 0x80000000


def key2f(k):
    u = (k & 0x7FFFFFFF) if k & 0x80000000 else (~k) & 0xFFFFFFFF
    return np.uint32(u).view(np.float32)


def ring_pass(ring, r):
    p1 = p2 = s1 = z = 0
    for key in ring:
        u = (key - r) & 0xFFFFFFFF
        v = (r - key) & 0xFFFFFFFF
        p2 = max(p2, min(p1, u)); p1 = max(p1, u); s1 = max(s1, v); z += min(u, 1)
    return p1, p2, s1, len(ring) - z


def rolling_median(x, w):
    n = len(x)
    out = np.full(n, np.nan, np.float32)
    k, even = w >> 1, (w & 1) == 0
    ring, pos, missing, valid = [], 0, 0, False
    r = cL = cG = 0
    known_lo = True
    for t in range(n):
        present = bool(np.isfinite(x[t]))
        kn = f2key(x[t]) if present else MISSING
        ko = None
        if len(ring) == w:
            ko = ring[pos]; ring[pos] = kn
            missing -= ko == MISSING
        else:
            ring.append(kn)
        pos = (pos + 1) % w
        missing += not present
        if len(ring) < w:
            continue
        if missing:
            valid = False
            continue
        if not valid:
            r, cL = 0, 0
            while True:
                p1, p2, s1, E = ring_pass(ring, r)
                if cL + E > k:
                    break
                cL += E; r = (r - s1) & 0xFFFFFFFF
            cG = w - cL - E
            hi = r; lo = r if cL <= k - 1 else (r + p1) & 0xFFFFFFFF
            valid, known_lo = True, True
        else:
            if known_lo:
                cL += (kn < r) - (ko < r)
            else:
                cG += (kn > r) - (ko > r)
            p1, p2, s1, E = ring_pass(ring, r)
            if known_lo:
                cG = w - cL - E
            else:
                cL = w - cG - E
            if cL > k:
                assert cL == k + 1                      # ranks move by at most one
                hi = (r + p1) & 0xFFFFFFFF; lo = (r + p2) & 0xFFFFFFFF
                cG = w - cL; known_lo = False; r = hi
            elif cL + E <= k:
                assert cL + E == k
                hi = (r - s1) & 0xFFFFFFFF; lo = r if E > 0 else (r + p1) & 0xFFFFFFFF
                cL += E; known_lo = True; r = hi
            else:
                hi = r; lo = r if cL <= k - 1 else (r + p1) & 0xFFFFFFFF
                known_lo = True
        out[t] = np.float32(0.5 * (float(key2f(lo)) + float(key2f(hi)))) if even else key2f(hi)
    return out


def _series(rng, n, mode):
    if mode == 0:
        x = rng.standard_normal(n)
    elif mode == 1:
        x = rng.integers(-2, 3, n).astype(float)                      # heavy ties, negatives, zeros
    elif mode == 2:
        x = np.where(rng.random(n) < 0.1, np.nan, rng.integers(0, 4, n))
    elif mode == 3:
        x = np.where(rng.random(n) < 0.03, np.nan, rng.standard_normal(n))
        x[rng.integers(0, n)] = np.inf; x[rng.integers(0, n)] = -np.inf
    else:
        x = np.where(rng.random(n) < 0.5, 0.0, -0.0)
    return x.astype(np.float32)


@pytest.mark.parametrize("mode", range(5))
def test_rank_tracking_median_equals_pandas(mode):
    rng = np.random.default_rng(100 + mode)
    for trial in range(14):
        w = int(rng.integers(1, 14))
        x = _series(rng, int(rng.integers(1, 120)), mode)
        want = pd.Series(x).rolling(w).median().to_numpy().astype(np.float32)
        np.testing.assert_array_equal(rolling_median(x, w), want, err_msg=f"mode {mode} window {w}")


def test_rank_tracking_median_long_even_window():
    rng = np.random.default_rng(7)
    x = rng.standard_normal(700).astype(np.float32)
    x[300:310] = np.nan
    want = pd.Series(x).rolling(144).median().to_numpy().astype(np.float32)
    np.testing.assert_array_equal(rolling_median(x, 144), want)
