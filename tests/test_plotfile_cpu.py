"""CPU: the AMReX-free plotfile reader / writer: write -> read round trip, and — where the
reference tree is mounted — byte identity of the re-written Level files with the reference's own fixture
(which its writer test demands of AMReX, src/writeplotfile.cpp:400)."""
import filecmp
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_PLT = "/root/reference/tests/plt00074"


def test_write_read_roundtrip(tmp_path):
    wc = importlib.import_module("wavelet-compression_b200")
    rng = np.random.default_rng(0)
    boxes = [((0, 0, 0), (15, 7, 11)), ((16, 8, 12), (31, 15, 15))]
    data = [rng.standard_normal((3, 12, 8, 16)), rng.standard_normal((3, 4, 8, 16))]
    wc.plotfile.write_level(str(tmp_path / "plt0"), 0, boxes, data, 3)
    lev = wc.plotfile.read_level(str(tmp_path / "plt0"), 0)
    assert lev.ncomp == 3 and len(lev.fabs) == 2
    for fab, (lo, hi), d in zip(lev.fabs, boxes, data):
        assert fab.lo == lo and fab.hi == hi and np.array_equal(fab.data, d)
    units = wc.plotfile.level_units(lev, [2, 0])
    assert len(units) == 4 and units[0][1] == (16, 8, 12) and np.array_equal(units[1][0], data[0][0])


def test_fold_minmax_reference_quirk():
    wc = importlib.import_module("wavelet-compression_b200")
    lo, hi = wc.modes.fold_minmax([-3.0, 5.0, -7.0, 1.0], [-1.0, 9.0, -2.0, 4.0], 2)
    assert lo == [-7.0, 1.0]
    assert hi[1] == 9.0 and hi[0] == float(np.finfo(np.float32).tiny)   # all-negative component keeps FLT_MIN


@pytest.mark.skipif(not os.path.isdir(REF_PLT), reason="reference fixtures not mounted")
def test_reference_fixture_rewritten_byte_identical(tmp_path):
    wc = importlib.import_module("wavelet-compression_b200")
    hdr = wc.plotfile.read_header(REF_PLT)
    assert hdr.names == ["temp", "pressure"] and hdr.finest_level == 1 and hdr.level_steps == [1200, 1500]
    assert hdr.domains[0] == ((0, 0, 0), (255, 511, 255))
    for level in (0, 1):
        lev = wc.plotfile.read_level(REF_PLT, level)
        assert [f.dims for f in lev.fabs] == [(16, 32, 64), (8, 4, 2)]
        assert float(lev.fabs[0].data[0, 0, 0, 0]) == 3902.39990234375 and float(lev.fabs[1].data[1, 0, 0, 0]) == 16.0
        out = str(tmp_path / "plt00074")
        wc.plotfile.write_level(out, level, [(f.lo, f.hi) for f in lev.fabs], [f.data for f in lev.fabs], lev.ncomp)
        for name in ("Cell_H", "Cell_D_00000"):
            assert filecmp.cmp(os.path.join(out, f"Level_{level}", name), os.path.join(REF_PLT, f"Level_{level}", name),
                               shallow=False), name
    wc.plotfile.write_header(str(tmp_path / "plt00074"), hdr)
    assert wc.plotfile.read_header(str(tmp_path / "plt00074")) == hdr
