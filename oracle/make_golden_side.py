"""TEST INFRASTRUCTURE.  Generates tests/golden/sidefiles_v1.npz and tests/golden/plt_fixtures.tar.xz in THIS container
(needs /root/reference):

  * the five .raw side files written by the reference's own code (oracle/_ref/libwcref_side.so = its unmodified
    src/readandwrite.cpp) for three runs: the doctest values of src/readandwrite.cpp:397-490, the BASELINE config-2
    run over the bundled plt00074 / plt00075 (levels 0-1, temp + pressure), and a ragged 3-timestep run;
  * a copy of the reference's TEST DATA tests/plt00074 and tests/plt00075 (AMReX plotfiles of constant boxes, the
    byte-identity fixture of src/writeplotfile.cpp:400) so that GPU-box tests can compare regenerated plotfiles with
    it where /root/reference does not exist.  Data only; no reference source is copied.

    python oracle/make_golden_side.py
"""
import ctypes as C
import io
import json
import os
import sys
import tarfile
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

CASES = {
    "doctest": dict(counts=[[1, 1], [1, 1]], locs=[[0, 14, 44]] * 4, dims=[[16, 32, 64]] * 4,
                    files=["../../../raw/plt00740", "../../../raw/plt07500"], min_level=0, max_level=1,
                    comps=["Temp", "pressure"], comp_idxs=[6, 25],
                    geom=[[0.6, 0.5, 0.4, 0.8, 0.9, 1.0]] * 2, ref=[2, 2, 2], times=["0.2219392", "0.3874982"],
                    steps=[[1200, 1500], [1800, 2000]], xyz=[256, 512, 256]),
    "config2": dict(counts=[[2, 2], [2, 2]], locs=[[0, 0, 0], [16, 32, 64]] * 4, dims=[[16, 32, 64], [8, 4, 2]] * 4,
                    files=["../tests/plt00074", "../tests/plt00075"], min_level=0, max_level=1,
                    comps=["temp", "pressure"], comp_idxs=[0, 1],
                    # the time lines exactly as the bundled Headers spell them (parsed to long double by operator>>)
                    geom=[[0.6, 0.5, 0.4, 0.8, 0.9, 1.0]] * 2, ref=[2, 0, 0], times=["0.2219392", "0.38749820000000001"],
                    steps=[[1200, 1500], [1800, 2000]], xyz=[256, 512, 256]),
    "ragged": dict(counts=[[3, 1, 0], [2, 2, 5], [1, 0, 4]],
                   locs=[[i, 2 * i, 100 + i] for i in range(18)], dims=[[8 + i, 4, 2 * i + 2] for i in range(18)],
                   files=["/data/run/plt00000", "/data/run/plt00010", "/data/run/plt00020"], min_level=1, max_level=3,
                   comps=["density", "Temp", "Y(H2O)"], comp_idxs=[0, 6, 13],
                   geom=[[0.0, -1.5, 1e-9, 3.25, 1.5, 0.1 + 0.2]] * 3, ref=[2, 4, 2],
                   times=["0", "1.0000000000000000000123e-5", "3.141592653589793238462643383279"],
                   steps=[[0, 0, 0], [10, 20, 40], [20, 40, 80]], xyz=[64, 32, 16]),
}


def main():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libwcref_side.so"))
    nc = C.c_int(0)
    assert lib.wcref_side_doctests(C.byref(nc)) == 0 and nc.value == 4, "the reference's own side-file doctests must pass"
    out = {"manifest": None}
    for name, c in CASES.items():
        T, L = len(c["counts"]), len(c["counts"][0])
        with tempfile.TemporaryDirectory() as d:
            arr = lambda v, t=np.int32: np.ascontiguousarray(np.asarray(v, t).reshape(-1))
            counts, locs, dims = arr(c["counts"]), arr(c["locs"]), arr(c["dims"])
            ci, geom, ref, steps = arr(c["comp_idxs"]), arr(c["geom"], np.float64), arr(c["ref"]), arr(c["steps"])
            p = lambda a: a.ctypes.data_as(C.c_void_p)
            rc = lib.wcref_side_write((d + "/").encode(), T, L, p(counts), p(locs), p(dims), "\n".join(c["files"]).encode(),
                                      c["min_level"], c["max_level"], "\n".join(c["comps"]).encode(), p(ci), len(c["comp_idxs"]),
                                      p(geom), p(ref), "\n".join(c["times"]).encode(), p(steps), *c["xyz"])
            assert rc == 0
            for f in ("runinfo.raw", "locations.raw", "dimensions.raw", "boxcounts.raw", "amrexinfo.raw"):
                out[f"{name}_{f}"] = np.frombuffer(open(os.path.join(d, f), "rb").read(), np.uint8)
    out["manifest"] = np.frombuffer(json.dumps(CASES).encode(), np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sidefiles_v1.npz"), **out)
    # the reference's plotfile fixtures (test data)
    buf = io.BytesIO()
    with tarfile.open(fileobj=buf, mode="w:xz") as tar:
        for plt in ("plt00074", "plt00075"):
            tar.add(os.path.join(REF, "tests", plt), arcname=plt)
    open(os.path.join(ROOT, "tests", "golden", "plt_fixtures.tar.xz"), "wb").write(buf.getvalue())
    print("wrote sidefiles_v1.npz and plt_fixtures.tar.xz", len(buf.getvalue()), "bytes")


if __name__ == "__main__":
    main()
