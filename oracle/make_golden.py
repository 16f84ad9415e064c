#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — generates tests/golden/golden_v1.npz by running the REFERENCE's own
code (oracle/_ref/libwcref.so, built from the unmodified sources under /root/reference/src) on a
deterministic corpus.  Run in the build container (the GPU box has no /root/reference):

    make -C oracle ref && python oracle/make_golden.py

Each case stores the input box, the dims, keep, and the reference's outputs for every row of
SURVEY.md §8(a): coefficients (F), pairs (T/M/P), serialized bytes (S, xz-decoded from the file
the reference wrote), reconstruction (U+I, read back by the reference's decompress()), RMSE (R).
The corpus contains the reference's doctest vectors (src/compressor.cpp:369-406,
src/calc-loss.cpp:68-86) and the bundled plotfile boxes (tests/plt00074, configs 1-2).
"""
import json
import lzma
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref  # noqa: E402

REF_ROOT = os.environ.get("WC_REF", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")

F999 = float(np.float32(0.999))     # what the CLI passes: a float widened to double (argparse.h:13)
F99 = float(np.float32(0.99))
F9999 = float(np.float32(0.9999))


def corpus():
    rng = np.random.default_rng(20261018)
    cases = []

    def add(name, dims, box, keep):
        X, Y, Z = dims
        box = np.asarray(box)
        assert box.size == X * Y * Z, name
        cases.append((name, dims, box.reshape(Z, Y, X), float(keep)))

    # --- the reference's own doctest vectors -------------------------------------------------
    b = np.full((16, 8, 4), 5.0, np.float32)            # Box3D test(4, 8, 16, 5.0f)
    for (x, y, z, v) in [(1, 2, 3, 8.5), (2, 5, 6, 5.44), (1, 1, 1, 3.3999932), (2, 2, 2, 3.19229),
                         (3, 5, 12, 199.39029)]:
        b[z, y, x] = np.float32(v)
    add("doctest_wavelet_decomposition", (4, 8, 16), b, F999)
    add("doctest_file_roundtrip_const", (4, 8, 16), np.full(512, 5.0, np.float32), 0.999)

    # --- smooth + noise, the three README keeps -------------------------------------------------
    def smooth(X, Y, Z, noise=1e-3):
        i, j, k = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
        f = 300 + 50 * np.sin(0.1 * i) * np.cos(0.07 * j) * np.sin(0.05 * k)
        f = f + noise * rng.standard_normal(f.shape)
        return np.ascontiguousarray(f.transpose(2, 1, 0)).astype(np.float32)

    for keep, tag in [(F99, "k99"), (F999, "k999"), (F9999, "k9999")]:
        add(f"smooth_8x8x8_{tag}", (8, 8, 8), smooth(8, 8, 8), keep)
    add("smooth_16x8x32_k999", (16, 8, 32), smooth(16, 8, 32), F999)
    add("smooth_6x10x14_k99", (6, 10, 14), smooth(6, 10, 14), F99)
    add("smooth_12x4x20_k9999", (12, 4, 20), smooth(12, 4, 20), F9999)
    add("smooth_32x32x32_k999", (32, 32, 32), smooth(32, 32, 32), F999)

    # --- odd / tiny / irregular dims (D7) ------------------------------------------------------------
    add("ramp_3x5x7", (3, 5, 7), np.arange(105, dtype=np.float32), F999)
    add("rand_3x5x7", (3, 5, 7), rng.standard_normal(105).astype(np.float32), F99)
    add("rand_5x4x6", (5, 4, 6), (10 + rng.standard_normal(120)).astype(np.float32), F999)
    add("rand_7x7x7", (7, 7, 7), (10 + rng.standard_normal(343)).astype(np.float32), F99)
    add("rand_9x1x3", (9, 1, 3), rng.standard_normal(27).astype(np.float32), F999)
    add("rand_1x8x8", (1, 8, 8), (5 + rng.standard_normal(64)).astype(np.float32), F999)
    add("one_1x1x1", (1, 1, 1), np.array([3.25], np.float32), F999)
    add("neg_1x1x1", (1, 1, 1), np.array([-3.25], np.float32), F999)
    add("cube_2x2x2", (2, 2, 2), np.arange(8, dtype=np.float32) - 3, F999)
    add("line_2x1x1", (2, 1, 1), np.array([1.0, -4.0], np.float32), F999)
    add("line_1x1x4", (1, 1, 4), np.array([1.0, 2.0, 3.0, 5.0], np.float32), F99)
    add("fixture_small_8x4x2", (8, 4, 2), np.full(64, 16.0, np.float32), F999)

    # --- threshold semantics (D3, D3') -------------------------------------------------------------
    z = np.zeros(64, np.float32); z[21] = -8.0
    add("negmax_4x4x4", (4, 4, 4), z, F999)
    z = np.zeros(64, np.float32); z[21] = 8.0
    add("posmax_4x4x4", (4, 4, 4), z, F999)
    z = np.zeros(64, np.float32); z[2] = 8.0; z[60] = -8.0
    add("tie_pos_first_4x4x4", (4, 4, 4), z, F999)
    z = np.zeros(64, np.float32); z[2] = -8.0; z[60] = 8.0
    add("tie_neg_first_4x4x4", (4, 4, 4), z, F999)
    add("all_zero_4x4x4", (4, 4, 4), np.zeros(64, np.float32), F999)
    add("keep_one_4x4x4", (4, 4, 4), rng.standard_normal(64).astype(np.float32), 1.0)
    add("keep_zero_4x4x4", (4, 4, 4), (3 + rng.standard_normal(64)).astype(np.float32), 0.0)
    add("keep_half_8x8x8", (8, 8, 8), (3 + rng.standard_normal(512)).astype(np.float32), 0.5)
    add("symmetric_8x8x8", (8, 8, 8), rng.standard_normal(512).astype(np.float32), F999)

    # --- special values ------------------------------------------------------------------------------
    z = rng.standard_normal(64).astype(np.float32); z[0:8] = np.nan
    add("nan_first_block_4x4x4", (4, 4, 4), z, F999)
    z = rng.standard_normal(64).astype(np.float32); z[37] = np.nan
    add("nan_inside_4x4x4", (4, 4, 4), z, F999)
    z = rng.standard_normal(64).astype(np.float32); z[11] = np.inf
    add("inf_4x4x4", (4, 4, 4), z, F999)
    z = rng.standard_normal(64).astype(np.float32); z[11] = -np.inf; z[40] = np.inf
    add("neg_inf_first_4x4x4", (4, 4, 4), z, F999)
    z = np.ldexp(rng.uniform(-1, 1, 512), rng.integers(-140, 120, 512)).astype(np.float32)
    add("wide_exponents_8x8x8", (8, 8, 8), z, F999)
    z = (rng.uniform(-1, 1, 64) * 1e-44).astype(np.float32)
    add("denormals_4x4x4", (4, 4, 4), z, F99)
    z = rng.uniform(-1, 1, 64).astype(np.float32) * np.float32(3e38)
    add("near_overflow_4x4x4", (4, 4, 4), z, F99)

    # --- float64 ingest (A1): values that do not round-trip through float32 ------------------------
    d = 1500.0 + 800.0 * rng.standard_normal((8, 8, 8)) + 1e-9 * rng.standard_normal((8, 8, 8))
    cases.append(("f64_ingest_8x8x8", (8, 8, 8), d.astype(np.float64), F999))
    d = rng.standard_normal((6, 4, 10)) * 50.0
    cases.append(("f64_ingest_sym_10x4x6", (10, 4, 6), d.astype(np.float64), F9999))
    return cases


def read_fab_file(level_dir):
    """Boxes of one plotfile level straight from Cell_H / Cell_D: returns
    [(dims, float64 array [ncomp][nz][ny][nx])]."""
    hdr = open(os.path.join(level_dir, "Cell_H")).read().split("\n")
    ncomp = int(hdr[2])
    m = re.match(r"\((\d+) \d+", hdr[4])
    nbox = int(m.group(1))
    boxes = []
    for line in hdr[5:5 + nbox]:
        lo, hi = re.findall(r"\((-?\d+),(-?\d+),(-?\d+)\)", line)[:2]
        boxes.append(tuple(int(h) - int(l) + 1 for l, h in zip(lo, hi)))
    fods = [l for l in hdr if l.startswith("FabOnDisk:")]
    out = []
    for dims, fod in zip(boxes, fods):
        _, fname, off = fod.split()
        with open(os.path.join(level_dir, fname), "rb") as f:
            f.seek(int(off))
            f.readline()  # the ASCII "FAB ((8, (64 11 52 ...\n" line
            n = dims[0] * dims[1] * dims[2]
            data = np.frombuffer(f.read(8 * n * ncomp), "<f8").reshape(ncomp, dims[2], dims[1], dims[0])
        out.append((dims, data.copy()))
    return out


def main():
    ref = Ref()
    fails, names, nassert = ref.run_doctests()
    assert fails == 0, "the reference's own doctests fail in this build"
    manifest = {"generator": "oracle/make_golden.py", "reference_doctests": names,
                "reference_doctest_assertions": nassert, "cases": []}
    arrays = {}

    def run_case(name, dims, box_in, keep):
        box32 = box_in.astype(np.float32)  # src/preprocess.cpp:78 when the input is float64
        with tempfile.TemporaryDirectory() as d:
            coef = ref.haar_forward(box32, dims)
            (runs, vals), = ref.compress(box32.reshape(1, -1), dims, keep, d)
            path = os.path.join(d, "compressed-wavelet-0-0-0-0.xz")
            ser = np.frombuffer(lzma.decompress(open(path, "rb").read()), np.uint8)
            recon, rdims = ref.decompress(path)
            assert rdims == tuple(dims)
            rmse = ref.rmse(box32, recon, dims)[0]
        i = len(manifest["cases"])
        manifest["cases"].append({"name": name, "dims": list(dims), "keep": keep,
                                  "in_dtype": str(box_in.dtype), "npairs": int(runs.size),
                                  "serialized_bytes": int(ser.size), "rmse": repr(float(rmse))})
        arrays[f"c{i}_in"] = box_in
        arrays[f"c{i}_coef"] = coef
        arrays[f"c{i}_runs"] = runs
        arrays[f"c{i}_vals"] = vals
        arrays[f"c{i}_ser"] = ser
        arrays[f"c{i}_recon"] = recon
        arrays[f"c{i}_rmse"] = np.array([rmse])

    for name, dims, box, keep in corpus():
        run_case(name, dims, box, keep)

    # bundled plotfiles: configs 1 and 2 (components temp, pressure)
    for plt in ("plt00074", "plt00075"):
        for lev in (0, 1):
            for bi, (dims, data) in enumerate(read_fab_file(os.path.join(REF_ROOT, "tests", plt, f"Level_{lev}"))):
                for ci, cname in enumerate(("temp", "pressure")):
                    for keep, tag in ((F999, "k999"), (F9999, "k9999")):
                        if plt == "plt00075" and tag == "k999":
                            continue
                        run_case(f"fixture_{plt}_L{lev}_b{bi}_{cname}_{tag}", dims, data[ci], keep)

    arrays["manifest"] = np.frombuffer(json.dumps(manifest).encode(), np.uint8)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **arrays)
    print(f"wrote {OUT}: {len(manifest['cases'])} cases, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
