"""CPU: the side files of a compression run (sidefiles.py) against bytes written by the reference's own code
(tests/golden/sidefiles_v1.npz, generated from the unmodified src/readandwrite.cpp by oracle/make_golden_side.py), the
Header the `-d` mode writes against the reference's bundled fixtures (tests/golden/plt_fixtures.tar.xz), and — where
the reference build is present — a round trip of this writer's files through the reference's readers and writers."""
import ctypes as C
import importlib
import json
import os
import sys
import tarfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SIDE_SO = os.path.join(ROOT, "oracle", "_ref", "libwcref_side.so")


@pytest.fixture(scope="module")
def wc():
    return importlib.import_module("wavelet-compression_b200")


@pytest.fixture(scope="module")
def golden_side():
    z = np.load(os.path.join(ROOT, "tests", "golden", "sidefiles_v1.npz"))
    return z, json.loads(bytes(z["manifest"]))


def _nest(case):
    """flat per-box lists -> data[t][level][box]"""
    locs, dims, k = [], [], 0
    for t in case["counts"]:
        locs.append([]); dims.append([])
        for c in t:
            locs[-1].append([tuple(v) for v in case["locs"][k:k + c]])
            dims[-1].append([tuple(v) for v in case["dims"][k:k + c]])
            k += c
    return locs, dims


def _write_case(sf, d, case):
    locs, dims = _nest(case)
    ri = sf.RunInfo(case["files"], case["min_level"], case["max_level"], case["comps"], case["comp_idxs"])
    ai = sf.AMReXInfo(case["geom"], case["ref"], case["times"], case["steps"], *case["xyz"])
    sf.write_all(d, ri, locs, dims, case["counts"], ai)
    return ri, ai, locs, dims


@pytest.mark.parametrize("name", ["doctest", "config2", "ragged"])
def test_writer_matches_reference_bytes_and_reader_parses_them(wc, golden_side, tmp_path, name):
    sf = wc.sidefiles
    z, cases = golden_side
    case = cases[name]
    d = str(tmp_path) + "/"
    ri, ai, locs, dims = _write_case(sf, d, case)
    for f in sf.NAMES:
        got, want = open(d + f, "rb").read(), z[f"{name}_{f}"].tobytes()
        if f == "amrexinfo.raw":
            got, want = sf.mask_long_double_padding(got), sf.mask_long_double_padding(want)
        assert got == want, (name, f)
    # the reader on the REFERENCE-written bytes
    g = str(tmp_path / "golden") + "/"
    os.makedirs(g)
    for f in sf.NAMES:
        open(g + f, "wb").write(z[f"{name}_{f}"].tobytes())
    r = sf.read_runinfo(g)
    assert (r.files, r.min_level, r.max_level, r.components, r.comp_idxs) == (ri.files, ri.min_level, ri.max_level, ri.components, ri.comp_idxs)
    nt, nl = len(r.files), r.max_level - r.min_level + 1
    counts = sf.read_box_counts(g, nt, nl)
    assert counts == case["counts"]
    assert sf.read_loc_dim(g, "locations.raw", counts) == locs and sf.read_loc_dim(g, "dimensions.raw", counts) == dims
    a = sf.read_amrexinfo(g)
    assert a.geomcellinfo == case["geom"] and a.ref_ratios == case["ref"] and a.level_steps == case["steps"]
    assert (a.xDim, a.yDim, a.zDim) == tuple(case["xyz"])
    assert [np.longdouble(t) for t in case["times"]] == list(a.true_times)      # long double precision survives


@pytest.mark.skipif(not os.path.exists(SIDE_SO), reason="oracle/_ref/libwcref_side.so not built (needs /root/reference)")
def test_reference_readers_accept_this_writers_files(wc, golden_side, tmp_path):
    lib = C.CDLL(SIDE_SO)
    nc = C.c_int(0)
    # 4 cases of src/readandwrite.cpp (+ the 5 of libwcref.so when that is loaded too: the shim registry is process-wide)
    assert lib.wcref_side_doctests(C.byref(nc)) == 0 and nc.value in (4, 9)
    sf = wc.sidefiles
    _, cases = golden_side
    for name, case in cases.items():
        a, b = str(tmp_path / name / "mine") + "/", str(tmp_path / name / "ref") + "/"
        os.makedirs(a); os.makedirs(b)
        _write_case(sf, a, case)
        rc = lib.wcref_side_rewrite(a.encode(), b.encode())
        assert rc == len(case["counts"]) * 1000 + len(case["counts"][0])
        for f in sf.NAMES:
            x, y = open(a + f, "rb").read(), open(b + f, "rb").read()
            if f == "amrexinfo.raw":
                x, y = sf.mask_long_double_padding(x), sf.mask_long_double_padding(y)
            assert x == y, (name, f)


def test_amrexinfo_from_headers_and_header_identity_with_the_reference_fixture(wc, tmp_path):
    """What src/preprocess.cpp:166-259 extracts from the Headers (quirks included) and, from that, the Header
    WriteMultiLevelPlotfile writes: byte-identical to the reference's bundled plt00074 / plt00075."""
    with tarfile.open(os.path.join(ROOT, "tests", "golden", "plt_fixtures.tar.xz")) as tar:
        tar.extractall(str(tmp_path), filter="data")
    plts = [str(tmp_path / p) for p in ("plt00074", "plt00075")]
    info = wc.sidefiles.amrexinfo_from_headers(plts, 2)
    assert info.ref_ratios == [2, 0, 0]                 # `dim` ints off a one-entry line: the reference's quirk, kept
    assert (info.xDim, info.yDim, info.zDim) == (256, 512, 256) and info.level_steps == [[1200, 1500], [1800, 2000]]
    assert info.geomcellinfo == [[0.6, 0.5, 0.4, 0.8, 0.9, 1.0]] * 2 and info.true_times[0] == "0.2219392"
    for t, plt in enumerate(plts):
        boxes = []
        for level in (0, 1):
            lev = wc.plotfile.read_level(plt, level)
            boxes.append([(f.lo, f.hi) for f in lev.fabs])
        out = str(tmp_path / "out" / os.path.basename(plt))
        wc.plotfile.write_header(out, wc.plotfile.header_from_amrexinfo(["temp", "pressure"], info, t, boxes))
        assert open(os.path.join(out, "Header"), "rb").read() == open(os.path.join(plt, "Header"), "rb").read()


def test_read_level_refuses_ghost_cells_and_mismatched_fab_headers(wc, tmp_path):
    rng = np.random.default_rng(0)
    boxes = [((0, 0, 0), (7, 3, 1))]
    wc.plotfile.write_level(str(tmp_path / "p"), 0, boxes, [rng.standard_normal((1, 2, 4, 8))], 1)
    wc.plotfile.read_level(str(tmp_path / "p"), 0)
    h = tmp_path / "p" / "Level_0" / "Cell_H"
    lines = h.read_text().split("\n")
    lines[3] = "2"
    h.write_text("\n".join(lines))
    with pytest.raises(ValueError, match="nghost"):
        wc.plotfile.read_level(str(tmp_path / "p"), 0)
    lines[3] = "0"
    lines[5] = "((0,0,0) (7,3,2) (0,0,0))"
    h.write_text("\n".join(lines))
    with pytest.raises(ValueError, match="FAB header box"):
        wc.plotfile.read_level(str(tmp_path / "p"), 0)
