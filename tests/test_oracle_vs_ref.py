"""CPU: pins the oracle restatement against the REFERENCE's own code compiled here
(oracle/_ref/libwcref.so).  Skipped where /root/reference was never available (the GPU box carries
the prebuilt library, so it normally runs there too)."""
import lzma
import os
import tempfile

import numpy as np
import pytest

from conftest import same_bits


def test_reference_doctests_pass_in_this_build(ref):
    fails, names, nassert = ref.run_doctests()
    assert fails == 0 and len(names) == 5 and nassert >= 10


def _check(oracle, ref, dims, box, keep):
    fo, fr = oracle.haar_forward(box, dims), ref.haar_forward(box, dims)
    assert same_bits(fo, fr)
    assert same_bits(oracle.haar_inverse(fo, dims), ref.haar_inverse(fr, dims))
    with tempfile.TemporaryDirectory() as d:
        (rr, rv), = ref.compress(box.reshape(1, -1), dims, keep, d)
        orr, ov, _ = oracle.compress_unit(box, dims, keep)
        assert same_bits(rr, orr) and same_bits(rv, ov)
        path = os.path.join(d, "compressed-wavelet-0-0-0-0.xz")
        assert lzma.decompress(open(path, "rb").read()) == oracle.packed_bytes(box, dims, keep).tobytes()
        db, dd = ref.decompress(path)
        ob = oracle.decompress_unit(orr, ov, dims)
        assert dd == tuple(dims) and same_bits(db, ob)
        a, b = ref.rmse(box, db, dims)[0], oracle.rmse(box, ob, dims)
        assert a == b or (np.isnan(a) and np.isnan(b))


@pytest.mark.parametrize("dims", [(4, 8, 16), (8, 4, 2), (3, 5, 7), (1, 1, 1), (2, 2, 2), (5, 4, 6),
                                  (6, 7, 2), (1, 8, 8), (9, 1, 3), (16, 16, 16)])
def test_random_boxes(oracle, ref, dims):
    rng = np.random.default_rng(hash(dims) % 2**32)
    n = dims[0] * dims[1] * dims[2]
    for keep in [float(np.float32(0.999)), float(np.float32(0.99)), 0.5, 1.0, 0.0]:
        _check(oracle, ref, dims, (300 + 50 * rng.standard_normal(n)).astype(np.float32), keep)
        _check(oracle, ref, dims, rng.standard_normal(n).astype(np.float32), keep)
        _check(oracle, ref, dims, np.ldexp(rng.uniform(-1, 1, n), rng.integers(-60, 60, n)).astype(np.float32), keep)


def test_special_values(oracle, ref):
    rng = np.random.default_rng(11)
    for dims in [(4, 4, 4), (3, 5, 7)]:
        n = dims[0] * dims[1] * dims[2]
        for trial in range(12):
            box = rng.standard_normal(n).astype(np.float32)
            idx = rng.integers(0, n, 3)
            box[idx[0]] = np.nan if trial % 3 == 0 else np.inf
            if trial % 2:
                box[idx[1]] = -np.inf
            if trial % 5 == 0:
                box[:8] = np.nan
            _check(oracle, ref, dims, box, 0.999)


def test_rle_and_serialize_primitives(oracle, ref):
    rng = np.random.default_rng(5)
    mask = rng.random(200) < 0.3
    flat = np.where(mask, rng.standard_normal(200) + 5, 0).astype(np.float32)
    rr, rv = ref.rle_encode(mask.astype(np.uint8), flat[mask])
    orr, ov = oracle.threshold_pack(flat, 1e-30)
    assert same_bits(rr, orr) and same_bits(rv, ov)
    assert same_bits(ref.rle_decode(rr, rv, 200), oracle.rle_decode(orr, ov, 200))
    assert same_bits(ref.serialize((4, 5, 10), 200, rr, rv), oracle.serialize((4, 5, 10), 200, orr, ov))
    assert ref.deserialize(ref.serialize((4, 5, 10), 200, rr, rv))[0] == (4, 5, 10)
