"""GPU: the -estimate / -c / -d drivers (modes.py) over AMReX-format plotfiles written by plotfile.py.
BASELINE config 1 is reproduced from the bundled-fixture payloads stored in the golden file: 4096 + 8 pairs,
RMSE 0, adjusted loss 0, predicted compressed size 0.09737 %."""
import lzma
import os

import numpy as np
import pytest

from conftest import same_bits, smooth_box

pytestmark = pytest.mark.gpu
F999 = float(np.float32(0.999))


def _fixture_plotfile(wc, golden, tmp, plt="plt00074"):
    """Rebuilds Level_0 / Level_1 of the reference's bundled plotfile from the golden payloads."""
    by = {c["name"]: i for i, c in enumerate(golden.cases)}
    boxes = [((0, 0, 0), (15, 31, 63)), ((16, 32, 64), (23, 35, 65))]
    out = os.path.join(tmp, plt)
    for level in (0, 1):
        data = []
        for b in range(2):
            comps = [golden.arrays(by[f"fixture_{plt}_L{level}_b{b}_{name}_k9999"])["in"] for name in ("temp", "pressure")]
            data.append(np.stack(comps))
        wc.plotfile.write_level(out, level, boxes, data, 2)
    hdr = wc.plotfile.Header("HyperCLaw-V1.1", ["temp", "pressure"], 3, 0.2219392, 1, [0.6, 0.5, 0.4], [0.8, 0.9, 1.0], [2],
                             [((0, 0, 0), (255, 511, 255)), ((0, 0, 0), (511, 1023, 511))], [1200, 1500],
                             [[0.00078125] * 3, [0.000390625] * 3], 0, 0,
                             [dict(level=l, ngrids=2, time=0.2219392, step=s, boxes_phys=[[(0.0, 1.0)] * 3] * 2, path=f"Level_{l}/Cell")
                              for l, s in ((0, 1200), (1, 1500))])
    wc.plotfile.write_header(out, hdr)
    return out


def test_baseline_config1_estimate(wc, ctx, golden, tmp_path):
    plt = _fixture_plotfile(wc, golden, str(tmp_path))
    assert os.path.getsize(os.path.join(plt, "Level_0", "Cell_D_00000")) == 525493
    est = wc.modes.estimate(plt, 0, ["temp"], F999, ctx=ctx)
    assert est["npairs"] == [4096, 8]
    assert est["components"]["temp"]["rmse"] == 0.0 and est["components"]["temp"]["adjusted_loss"] == 0.0
    assert est["components"]["temp"]["max"] == np.float32(3902.4) and est["components"]["temp"]["min"] == 16.0
    assert abs(est["compressed_percent"] - 100.0 * 256 / 262913.5) < 1e-9       # 168 B + 88 B of .xz


def test_baseline_config2_roundtrip(wc, ctx, golden, tmp_path):
    """-c then -d on plt00074..plt00075, both levels, components temp + pressure, keep = 0.9999f: 16 units,
    4096 / 8 pairs each, regenerated Level files byte-identical to the inputs (constant boxes)."""
    import filecmp
    plts = [_fixture_plotfile(wc, golden, str(tmp_path / "in"), p) for p in ("plt00074", "plt00075")]
    cdir = str(tmp_path / "compressed")
    man = wc.modes.compress_run(plts, [0, 1], ["temp", "pressure"], float(np.float32(0.9999)), cdir, ctx=ctx)
    files = sorted(f for f in os.listdir(cdir) if f.endswith(".xz"))
    assert len(files) == 16
    for f in files:
        p = wc.PackedUnit.deserialize(lzma.decompress(open(os.path.join(cdir, f), "rb").read()))
        assert p.npairs == (4096 if p.dims == (16, 32, 64) else 8)
    wc.modes.decompress_run(cdir, str(tmp_path / "out"), ctx=ctx)
    for p in ("plt00074", "plt00075"):
        for level in (0, 1):
            for name in ("Cell_H", "Cell_D_00000"):
                assert filecmp.cmp(os.path.join(str(tmp_path / "out"), p, f"Level_{level}", name),
                                   os.path.join(str(tmp_path / "in"), p, f"Level_{level}", name), shallow=False)


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
        want = float(np.sum(per[name]) / len(per[name]))
        assert abs(est["components"][name]["rmse"] - want) <= 1e-12 * want
        assert est["components"][name]["min"] == lo[name] and est["components"][name]["max"] == hi[name]
        assert abs(est["components"][name]["adjusted_loss"] - want / (hi[name] - lo[name])) <= 1e-12 * want
