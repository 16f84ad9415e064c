"""GPU: the -estimate / -c / -d drivers (modes.py) over AMReX-format plotfiles.
BASELINE config 1 is reproduced on the reference's own bundled plotfile (tests/golden/plt_fixtures.tar.xz, a copy of
its test data): 4096 + 8 pairs, RMSE 0, adjusted loss 0, predicted compressed size 0.09737 %.
BASELINE config 2 (-c then -d on plt00074..plt00075) must regenerate the reference's fixture DIRECTORIES byte for byte —
Header, Cell_H and Cell_D — which is what the reference's writer test demands of AMReX (src/writeplotfile.cpp:400)."""
import filecmp
import lzma
import os
import tarfile

import numpy as np
import pytest

from conftest import same_bits, smooth_box

pytestmark = pytest.mark.gpu
F999 = float(np.float32(0.999))


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fixtures(tmp):
    """The reference's bundled plotfiles plt00074 / plt00075 (copied test data, oracle/make_golden_side.py)."""
    with tarfile.open(os.path.join(ROOT, "tests", "golden", "plt_fixtures.tar.xz")) as tar:
        tar.extractall(tmp, filter="data")
    return [os.path.join(tmp, p) for p in ("plt00074", "plt00075")]


def _same_tree(a, b):
    names = []
    for dp, _, files in os.walk(a):
        for f in files:
            rel = os.path.relpath(os.path.join(dp, f), a)
            names.append(rel)
            assert filecmp.cmp(os.path.join(a, rel), os.path.join(b, rel), shallow=False), rel
    return sorted(names)


def test_baseline_config1_estimate(wc, ctx, tmp_path):
    plt = _fixtures(str(tmp_path))[0]
    assert os.path.getsize(os.path.join(plt, "Level_0", "Cell_D_00000")) == 525493
    est = wc.modes.estimate(plt, 0, ["temp"], F999, ctx=ctx)
    assert est["npairs"] == [4096, 8] and est["need32"] == [False, False]
    assert est["components"]["temp"]["rmse"] == 0.0 and est["components"]["temp"]["adjusted_loss"] == 0.0
    assert est["components"]["temp"]["max"] == np.float32(3902.4) and est["components"]["temp"]["min"] == 16.0
    assert abs(est["compressed_percent"] - 100.0 * 256 / 262913.5) < 1e-9       # 168 B + 88 B of .xz
    # device-resident: the raw float64 slabs cross the bus ONCE (plus a few KB of descriptor tables)
    assert est["input_bytes"] == 8 * (32768 + 64)
    assert est["input_bytes"] <= est["h2d_bytes"] <= est["input_bytes"] + 4096


def test_baseline_config2_roundtrip_regenerates_the_reference_fixture(wc, ctx, tmp_path):
    """-c then -d on plt00074..plt00075, both levels, components temp + pressure, keep = 0.9999f: 16 units, 4096 / 8 pairs
    each; the five side files are what the reference's own writers produce for this run (golden, masked long double
    padding); the regenerated plotfile directories are byte-identical to the reference's fixtures, Header included."""
    plts = _fixtures(str(tmp_path / "in"))
    cdir = str(tmp_path / "compressed")
    res = wc.modes.compress_run(plts, 0, 1, ["temp", "pressure"], float(np.float32(0.9999)), cdir, ctx=ctx)
    assert res["units"] == 16 and not any(res["need32"])
    files = sorted(f for f in os.listdir(cdir) if f.endswith(".xz"))
    assert len(files) == 16
    for f in files:
        p = wc.PackedUnit.deserialize(lzma.decompress(open(os.path.join(cdir, f), "rb").read()))
        assert p.npairs == (4096 if p.dims == (16, 32, 64) else 8)
    side = np.load(os.path.join(ROOT, "tests", "golden", "sidefiles_v1.npz"))
    for name in wc.sidefiles.NAMES:
        got = open(os.path.join(cdir, name), "rb").read()
        want = side[f"config2_{name}"].tobytes()
        if name == "runinfo.raw":        # the golden run names its inputs ../tests/plt0007x; only the paths differ
            ri = wc.sidefiles.read_runinfo(cdir + "/")
            assert [os.path.basename(f) for f in ri.files] == ["plt00074", "plt00075"]
            assert (ri.min_level, ri.max_level, ri.components, ri.comp_idxs) == (0, 1, ["temp", "pressure"], [0, 1])
        elif name == "amrexinfo.raw":
            assert wc.sidefiles.mask_long_double_padding(got) == wc.sidefiles.mask_long_double_padding(want)
        else:
            assert got == want, name
    wc.modes.decompress_run(cdir, str(tmp_path / "out"), ctx=ctx)
    for p in ("plt00074", "plt00075"):
        names = _same_tree(os.path.join(str(tmp_path / "in"), p), os.path.join(str(tmp_path / "out"), p))
        assert names == ["Header", "Level_0/Cell_D_00000", "Level_0/Cell_H", "Level_1/Cell_D_00000", "Level_1/Cell_H"]


def test_estimate_matches_oracle_on_nontrivial_plotfile(wc, ctx, oracle, tmp_path):
    rng = np.random.default_rng(31)
    boxes = [((0, 0, 0), (31, 31, 31)), ((32, 0, 0), (63, 31, 31)), ((0, 32, 0), (15, 47, 23))]
    data = [np.stack([smooth_box(tuple(h - l + 1 for l, h in zip(lo, hi)), rng, dtype=np.float64, sym=(c == 2)) for c in range(3)])
            for lo, hi in boxes]
    plt = str(tmp_path / "plt00010")
    wc.plotfile.write_level(plt, 0, boxes, data, 3)
    wc.plotfile.write_header(plt, wc.plotfile.Header("HyperCLaw-V1.1", ["density", "Temp", "x_velocity"], 3, 0.0, 0, [0.0] * 3,
                                                     [1.0] * 3, [], [((0, 0, 0), (63, 47, 31))], [10], [[1 / 64] * 3], 0, 0,
                                                     [dict(level=0, ngrids=3, time=0.0, step=10, boxes_phys=[[(0.0, 1.0)] * 3] * 3,
                                                           path="Level_0/Cell")]))
    comps = ["Temp", "x_velocity"]
    est = wc.modes.estimate(plt, 0, comps, F999, ctx=ctx)
    xz_total, k = 0, 0
    per = {c: [] for c in comps}
    lo = {c: np.inf for c in comps}
    hi = {c: -np.inf for c in comps}
    for b, (blo, bhi) in enumerate(boxes):
        dims = tuple(h - l + 1 for l, h in zip(blo, bhi))
        for name, ci in zip(comps, (1, 2)):
            box = data[b][ci]
            runs, vals, _ = oracle.compress_unit(box, dims, F999)
            assert est["npairs"][k] == runs.size
            k += 1
            xz_total += len(lzma.compress(oracle.packed_bytes(box, dims, F999).tobytes(), format=lzma.FORMAT_XZ,
                                          check=lzma.CHECK_CRC64, preset=6))
            b32 = box.astype(np.float32)
            per[name].append(oracle.rmse(b32, oracle.decompress_unit(runs, vals, dims), dims))
            lo[name], hi[name] = min(lo[name], float(b32.min())), max(hi[name], float(b32.max()))
    ldir = os.path.join(plt, "Level_0")
    raw = sum(os.path.getsize(os.path.join(ldir, f)) for f in os.listdir(ldir)) / 3 * 2
    assert abs(est["compressed_percent"] - xz_total / raw * 100) < 1e-9
    for name in comps:
        want = wc.modes.sequential_mean(per[name])                      # std::accumulate, left to right
        assert abs(est["components"][name]["rmse"] - want) <= 1e-12 * want
        assert est["components"][name]["min"] == lo[name] and est["components"][name]["max"] == hi[name]
        rng32 = float(np.float32(hi[name]) - np.float32(lo[name]))      # max_values[c] - min_values[c] in float, src/modes.cpp:289
        assert abs(est["components"][name]["adjusted_loss"] - want / rng32) <= 1e-12 * want / rng32
    assert est["input_bytes"] <= est["h2d_bytes"] <= est["input_bytes"] + 4096
