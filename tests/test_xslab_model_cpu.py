"""CPU: the decomposition the x-slab kernels (csrc/wc_xslab.cu) rest on, restated in numpy and checked
against the oracle — no GPU, no product code on the path (the kernels themselves are tested in tests/test_gpu_xslab.py).

Claims pinned here, for boxes of any shape (odd dimensions included):
  1. the coefficients of an x-slab (block columns [a0, a1), plus the trailing plane x = X-1 for the last slab of an odd X) are
     whole planes i' of the flat order f = (i'*Y + j')*Z + k' (src/compressor.cpp:178-181): the low planes [a0, a1) and the
     high planes [hx+a0, hx+a1) (+ plane X-1) — two contiguous flat ranges that depend on the slab's own input cells only;
  2. a slab's pairs are placed in the unit's ordered pair list by four numbers per slab (kept count and last kept flat index of
     either range): low ranges of slabs 0..S-1 first, then the high ranges;
  3. the decoder needs, per slab, only the plane table tab[i'] = (first pair at/after flat index i'*Y*Z, flat index of the pair
     before it) and never reads the trailing planes / rows / columns of an odd axis: the inverse leaves those cells at +0
     (src/decompressor.cpp:99-108)."""
import numpy as np
import pytest

from conftest import same_bits, smooth_box

SHAPES = [(3, 5, 7), (7, 5, 3), (1, 1, 1), (1, 6, 4), (9, 1, 2), (31, 17, 9), (12, 10, 14), (13, 8, 6), (20, 9, 11)]


def slabs_of(X, S):
    hx = X // 2
    na = -(-hx // S) if hx else 0
    out = []
    for r in range(S):
        a0 = min(hx, r * na)
        a1 = min(hx, a0 + na)
        own1 = bool(X & 1) and r == S - 1
        out.append((a0, a1, own1))
    return hx, out


@pytest.mark.parametrize("dims", SHAPES)
@pytest.mark.parametrize("S", [1, 2, 4])
def test_xslab_decomposition_matches_the_oracle(oracle, dims, S):
    X, Y, Z = dims
    YZ = Y * Z
    rng = np.random.default_rng(X * 10007 + Y * 101 + Z + S)
    box = smooth_box(dims, rng, sym=True).astype(np.float32)              # (Z, Y, X)
    keep = float(np.float32(0.9))
    flat = oracle.haar_forward(box, dims)
    runs, vals, _ = oracle.compress_unit(box, dims, keep)
    hx, slabs = slabs_of(X, S)

    # 1. a slab's coefficient planes depend on its own cells only: transform a box that is zero outside the slab's x range
    for a0, a1, own1 in slabs:
        xs = list(range(2 * a0, 2 * a1)) + ([X - 1] if own1 else [])
        if not xs:
            continue
        masked = np.zeros_like(box)
        masked[:, :, xs] = box[:, :, xs]
        fm = oracle.haar_forward(masked, dims).reshape(X, YZ)
        planes = list(range(a0, a1)) + list(range(hx + a0, hx + a1)) + ([X - 1] if own1 else [])
        assert same_bits(fm[planes], flat.reshape(X, YZ)[planes]), (dims, S, a0)
        others = [p for p in range(X) if p not in planes]
        assert not fm[others].any()                                       # and the slab contributes to no other plane

    # 2. ordered packing from per-range counts: rebuild the unit's pair list slab by slab
    thresh, _ = oracle.select_threshold(flat, keep)
    kept = np.flatnonzero(np.abs(flat.astype(np.float64)) > thresh)
    assert kept.size == runs.size
    ranges = [(a0 * YZ, a1 * YZ) for a0, a1, _ in slabs] + \
             [((hx + a0) * YZ, (hx + a1 + (1 if own1 else 0)) * YZ) for a0, a1, own1 in slabs]
    pos, prev, got_runs, got_vals = 0, -1, [], []
    for f0, f1 in ranges:                                                 # low ranges of every slab, then the high ranges
        k = kept[(kept >= f0) & (kept < f1)]
        for f in k:
            got_runs.append(f - prev - 1)
            got_vals.append(flat[f])
            prev = f
        pos += k.size
    assert pos == runs.size and same_bits(np.array(got_runs, np.int32), runs) and same_bits(np.array(got_vals, np.float32), vals)

    # 3. decode per slab from the plane table, never touching the trailing planes of odd axes
    fidx = np.cumsum(runs.astype(np.int64) + 1) - 1                       # flat index of every pair
    tab_first = np.searchsorted(fidx, np.arange(X + 1) * YZ, side="left")
    want = oracle.decompress_unit(runs, vals, dims)                       # (Z, Y, X)
    got = np.full_like(want, 7.0)
    hy, hz = Y // 2, Z // 2
    for a0, a1, own1 in slabs:
        nl = a1 - a0
        C = np.zeros((2 * nl, Y, Z), np.float32)
        for base, lo in ((0, a0), (nl, hx + a0)):                         # the slab's two pair ranges
            p0, p1 = tab_first[lo], tab_first[lo + nl]
            for p in range(p0, p1):
                d = fidx[p] - lo * YZ
                C[base + d // YZ].reshape(-1)[d % YZ] = vals[p]
        for al in range(nl):
            for b in range(hy):
                for c in range(hz):
                    v = np.array([C[(nl if sx else 0) + al, (hy if sy else 0) + b, (hz if sz else 0) + c]
                                  for sz in (0, 1) for sy in (0, 1) for sx in (0, 1)], np.float32)
                    # inverse, X then Y then Z (src/decompressor.cpp:90-156): float32 add / sub
                    for q in range(4):
                        v[2 * q], v[2 * q + 1] = np.float32(v[2 * q] + v[2 * q + 1]), np.float32(v[2 * q] - v[2 * q + 1])
                    for zi in range(2):
                        for xi in range(2):
                            i0, i1 = zi * 4 + xi, zi * 4 + 2 + xi
                            v[i0], v[i1] = np.float32(v[i0] + v[i1]), np.float32(v[i0] - v[i1])
                    for q in range(4):
                        v[q], v[4 + q] = np.float32(v[q] + v[4 + q]), np.float32(v[q] - v[4 + q])
                    for o in range(8):
                        got[2 * c + (o >> 2), 2 * b + ((o >> 1) & 1), 2 * (a0 + al) + (o & 1)] = v[o]
        xs = list(range(2 * a0, 2 * a1)) + ([X - 1] if own1 else [])
        if Z & 1:
            got[Z - 1, :, xs] = 0.0
        if Y & 1:
            got[:, Y - 1, xs] = 0.0
        if own1:
            got[:, :, X - 1] = 0.0
    assert same_bits(got, want), (dims, S)
