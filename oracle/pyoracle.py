"""TEST INFRASTRUCTURE ONLY — ctypes front-ends for the parity oracle.

  Oracle  -> oracle/libwcoracle.so      the plain-C restatement (oracle/wc_oracle.c)
  Ref     -> oracle/_ref/libwcref.so    the reference's own unmodified sources (oracle/ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (wavelet-compression_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libwcoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libwcref.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(ref: bool | None = None) -> None:
    """Compile the oracle (always) and oracle/_ref (when the reference tree is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    have_ref = os.path.isdir(os.environ.get("WC_REF", "/root/reference") + "/src")
    if ref is None:
        ref = have_ref
    if ref and have_ref:
        subprocess.check_call(
            ["make", "-s", "-C", HERE, "ref", "REF=" + os.environ.get("WC_REF", "/root/reference")])
        # the drop-in test links libwcgpu.so: only once the product library has been built
        if os.path.exists(os.path.join(HERE, "..", "wavelet-compression_b200", "libwcgpu.so")):
            subprocess.check_call(
                ["make", "-s", "-C", HERE, "dropin", "REF=" + os.environ.get("WC_REF", "/root/reference")])


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Oracle:
    """The C restatement.  Boxes are numpy arrays indexed [k][j][i] (x fastest in memory), i.e.
    shape (Z, Y, X); `dims` is always given as (X, Y, Z) like the reference's Grid3D."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        L.wco_narrow_f64.argtypes = [_f64p, C.c_long, _f32p]
        L.wco_haar_forward.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p]
        L.wco_select_threshold.argtypes = [_f32p, C.c_long, C.c_double, C.POINTER(C.c_long)]
        L.wco_select_threshold.restype = C.c_double
        L.wco_threshold_pack.argtypes = [_f32p, C.c_long, C.c_double, _i32p, _f32p]
        L.wco_threshold_pack.restype = C.c_long
        L.wco_serialize.argtypes = [_i32p, C.c_int32, _i32p, _f32p, C.c_int32, _u8p]
        L.wco_serialize.restype = C.c_long
        L.wco_deserialize.argtypes = [_u8p, _i32p, _i32p, _i32p, _f32p, C.c_long]
        L.wco_deserialize.restype = C.c_long
        L.wco_rle_decode.argtypes = [_i32p, _f32p, C.c_long, C.c_long, _f32p]
        L.wco_haar_inverse.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p]
        L.wco_rmse.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.wco_rmse.restype = C.c_double
        L.wco_adj_loss.argtypes = [C.c_double, C.c_double]
        L.wco_adj_loss.restype = C.c_double
        L.wco_compress_unit.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_double, _i32p, _f32p,
                                        C.POINTER(C.c_double)]
        L.wco_compress_unit.restype = C.c_long
        L.wco_compress_unit_f64.argtypes = [_f64p, C.c_int, C.c_int, C.c_int, C.c_double, _i32p,
                                            _f32p, C.POINTER(C.c_double)]
        L.wco_compress_unit_f64.restype = C.c_long
        L.wco_decompress_unit.argtypes = [_i32p, _f32p, C.c_long, C.c_int, C.c_int, C.c_int,
                                          C.c_long, _f32p]
        L.wco_select_threshold_global.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_long),
                                                  C.c_int, C.c_double]
        L.wco_select_threshold_global.restype = C.c_double

    # -- rows of SURVEY.md §8(a) --------------------------------------------------------------
    def narrow(self, a64):
        a64 = _c(a64, np.float64)
        out = np.empty(a64.shape, np.float32)
        self.lib.wco_narrow_f64(a64.reshape(-1), a64.size, out.reshape(-1))
        return out

    def haar_forward(self, box, dims):
        X, Y, Z = dims
        box = _c(box, np.float32).reshape(-1)
        assert box.size == X * Y * Z
        flat = np.empty(X * Y * Z, np.float32)
        self.lib.wco_haar_forward(box, X, Y, Z, flat)
        return flat

    def select_threshold(self, flat, keep):
        flat = _c(flat, np.float32).reshape(-1)
        am = C.c_long(-1)
        t = self.lib.wco_select_threshold(flat, flat.size, float(keep), C.byref(am))
        return t, am.value

    def threshold_pack(self, flat, thresh):
        flat = _c(flat, np.float32).reshape(-1)
        runs = np.empty(max(flat.size, 1), np.int32)
        vals = np.empty(max(flat.size, 1), np.float32)
        k = self.lib.wco_threshold_pack(flat, flat.size, float(thresh), runs, vals)
        return runs[:k].copy(), vals[:k].copy()

    def serialize(self, dims, ncoef, runs, vals):
        runs = _c(runs, np.int32)
        vals = _c(vals, np.float32)
        k = runs.size
        out = np.empty(20 + 8 * k, np.uint8)
        n = self.lib.wco_serialize(np.asarray(dims, np.int32), int(ncoef),
                                   runs if k else np.zeros(1, np.int32),
                                   vals if k else np.zeros(1, np.float32), k, out)
        assert n == out.size
        return out

    def deserialize(self, buf):
        buf = _c(buf, np.uint8)
        k_guess = (buf.size - 20) // 8
        shape = np.zeros(3, np.int32)
        ncoef = np.zeros(1, np.int32)
        runs = np.empty(max(k_guess, 1), np.int32)
        vals = np.empty(max(k_guess, 1), np.float32)
        k = self.lib.wco_deserialize(buf, shape, ncoef, runs, vals, k_guess)
        assert k >= 0
        return tuple(int(s) for s in shape), int(ncoef[0]), runs[:k].copy(), vals[:k].copy()

    def rle_decode(self, runs, vals, total):
        runs = _c(runs, np.int32)
        vals = _c(vals, np.float32)
        out = np.empty(max(total, 1), np.float32)
        self.lib.wco_rle_decode(runs if runs.size else np.zeros(1, np.int32),
                                vals if vals.size else np.zeros(1, np.float32), runs.size, total,
                                out)
        return out[:total]

    def haar_inverse(self, flat, dims):
        X, Y, Z = dims
        flat = _c(flat, np.float32).reshape(-1)
        assert flat.size == X * Y * Z
        box = np.empty(X * Y * Z, np.float32)
        self.lib.wco_haar_inverse(flat, X, Y, Z, box)
        return box.reshape(Z, Y, X)

    def rmse(self, a, b, dims):
        X, Y, Z = dims
        return self.lib.wco_rmse(_c(a, np.float32).reshape(-1), _c(b, np.float32).reshape(-1), X,
                                 Y, Z)

    def adj_loss(self, rmse, rng):
        return self.lib.wco_adj_loss(rmse, rng)

    def compress_unit(self, box, dims, keep):
        """F->T->M->P.  box float32 or float64 (the latter is narrowed first, A1).
        Returns (runs, vals, thresh)."""
        X, Y, Z = dims
        n = X * Y * Z
        runs = np.empty(max(n, 1), np.int32)
        vals = np.empty(max(n, 1), np.float32)
        th = C.c_double(0.0)
        box = np.asarray(box)
        if box.dtype == np.float64:
            k = self.lib.wco_compress_unit_f64(_c(box, np.float64).reshape(-1), X, Y, Z,
                                               float(keep), runs, vals, C.byref(th))
        else:
            k = self.lib.wco_compress_unit(_c(box, np.float32).reshape(-1), X, Y, Z, float(keep),
                                           runs, vals, C.byref(th))
        return runs[:k].copy(), vals[:k].copy(), th.value

    def decompress_unit(self, runs, vals, dims, ncoef=None):
        X, Y, Z = dims
        runs = _c(runs, np.int32)
        vals = _c(vals, np.float32)
        box = np.empty(max(X * Y * Z, 1), np.float32)
        self.lib.wco_decompress_unit(runs if runs.size else np.zeros(1, np.int32),
                                     vals if vals.size else np.zeros(1, np.float32), runs.size, X,
                                     Y, Z, X * Y * Z if ncoef is None else ncoef, box)
        return box[:X * Y * Z].reshape(Z, Y, X)

    def packed_bytes(self, box, dims, keep):
        """The 20+8K-byte pre-LZMA buffer of one unit (row S)."""
        runs, vals, _ = self.compress_unit(box, dims, keep)
        return self.serialize(dims, dims[0] * dims[1] * dims[2], runs, vals)

    def select_threshold_global(self, flats, keep):
        flats = [_c(f, np.float32).reshape(-1) for f in flats]
        ptrs = (C.c_void_p * len(flats))(*[f.ctypes.data for f in flats])
        ns = (C.c_long * len(flats))(*[f.size for f in flats])
        return self.lib.wco_select_threshold_global(ptrs, ns, len(flats), float(keep))


def quantile_threshold(flats, keep):
    """EXTENSION oracle (no counterpart in the reference: parity UNPINNED, defined here): the threshold of
    WC_THRESH_QUANTILE over the concatenation of `flats` (one array: per-unit mode).  n counts every coefficient, NaNs
    included; Kt = n - floor(keep * n); threshold = the magnitude of rank Kt in descending order (NaNs last), or -1
    (keep every non-NaN coefficient) when fewer than Kt + 1 coefficients are comparable.  Mask: |c| > threshold."""
    a = np.abs(np.concatenate([np.asarray(f, np.float32).reshape(-1) for f in flats])) if len(flats) else np.zeros(0, np.float32)
    n = a.size
    kt = n - min(int(np.floor(float(keep) * n)), n)
    valid = a[~np.isnan(a)]
    if kt >= valid.size:
        return -1.0
    return float(np.partition(valid, valid.size - 1 - kt)[valid.size - 1 - kt])


class Ref:
    """The reference's own code (oracle/_ref/libwcref.so).  Same conventions as Oracle."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (build with `make -C oracle ref` where /root/reference exists)")
        L = self.lib = C.CDLL(path)
        L.wcref_run_doctests.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.wcref_doctest_name.argtypes = [C.c_int]
        L.wcref_doctest_name.restype = C.c_char_p
        L.wcref_haar_forward.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p]
        L.wcref_haar_inverse.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _f32p]
        L.wcref_rle_encode.argtypes = [_u8p, C.c_int, _f32p, C.c_int, _i32p, _f32p]
        L.wcref_rle_decode.argtypes = [_i32p, _f32p, C.c_int, C.c_int, _f32p]
        L.wcref_serialize.argtypes = [_i32p, C.c_int, _i32p, _f32p, C.c_int, _u8p, C.c_long]
        L.wcref_serialize.restype = C.c_long
        L.wcref_deserialize.argtypes = [_u8p, C.c_long, _i32p, _i32p, _i32p, _f32p, C.c_int]
        L.wcref_compress.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                     C.c_int, C.c_int, _i32p, C.c_char_p, _i32p, C.c_void_p,
                                     C.c_void_p]
        L.wcref_decompress.argtypes = [C.c_char_p, _f32p, C.c_long, _i32p]
        L.wcref_decompress.restype = C.c_long
        L.wcref_rmse.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _f64p]
        L.wcref_adj_loss.argtypes = [C.c_double, C.c_double]
        L.wcref_adj_loss.restype = C.c_double
        L.wcref_set_lzma_mode.argtypes = [C.c_int]
        L.wcref_lzma_seconds.restype = C.c_double

    def run_doctests(self):
        nc, na = C.c_int(0), C.c_int(0)
        fails = self.lib.wcref_run_doctests(C.byref(nc), C.byref(na))
        names = [self.lib.wcref_doctest_name(i).decode() for i in range(nc.value)]
        return fails, names, na.value

    def haar_forward(self, box, dims):
        X, Y, Z = dims
        flat = np.empty(max(X * Y * Z, 1), np.float32)
        self.lib.wcref_haar_forward(_c(box, np.float32).reshape(-1), X, Y, Z, flat)
        return flat[:X * Y * Z]

    def haar_inverse(self, flat, dims):
        X, Y, Z = dims
        box = np.empty(max(X * Y * Z, 1), np.float32)
        self.lib.wcref_haar_inverse(_c(flat, np.float32).reshape(-1), X, Y, Z, box)
        return box[:X * Y * Z].reshape(Z, Y, X)

    def rle_encode(self, mask, values):
        mask = _c(mask, np.uint8)
        values = _c(values, np.float32)
        runs = np.empty(max(mask.size, 1), np.int32)
        vals = np.empty(max(mask.size, 1), np.float32)
        k = self.lib.wcref_rle_encode(mask, mask.size, values if values.size else np.zeros(1, np.float32),
                                      values.size, runs, vals)
        return runs[:k].copy(), vals[:k].copy()

    def rle_decode(self, runs, vals, total):
        runs = _c(runs, np.int32)
        vals = _c(vals, np.float32)
        out = np.empty(max(total, 1), np.float32)
        self.lib.wcref_rle_decode(runs if runs.size else np.zeros(1, np.int32),
                                  vals if vals.size else np.zeros(1, np.float32), runs.size, total,
                                  out)
        return out[:total]

    def serialize(self, dims, ncoef, runs, vals):
        runs = _c(runs, np.int32)
        vals = _c(vals, np.float32)
        out = np.empty(20 + 8 * runs.size, np.uint8)
        n = self.lib.wcref_serialize(np.asarray(dims, np.int32), int(ncoef),
                                     runs if runs.size else np.zeros(1, np.int32),
                                     vals if vals.size else np.zeros(1, np.float32), runs.size,
                                     out, out.size)
        assert n == out.size, n
        return out

    def deserialize(self, buf):
        buf = _c(buf, np.uint8)
        cap = max((buf.size - 20) // 8, 1)
        shape = np.zeros(3, np.int32)
        ncoef = np.zeros(1, np.int32)
        runs = np.empty(cap, np.int32)
        vals = np.empty(cap, np.float32)
        k = self.lib.wcref_deserialize(buf, buf.size, shape, ncoef, runs, vals, cap)
        assert k >= 0
        return tuple(int(s) for s in shape), int(ncoef[0]), runs[:k].copy(), vals[:k].copy()

    def compress(self, boxes, dims, keep, out_dir, t=0, lev=0, box_idx=0, comp_ids=None,
                 want_pairs=True):
        """reference compress(multiBox3D&, ...): boxes float32 [ncomp][Z][Y][X].  Writes the .xz
        files into out_dir.  Returns a list of (runs, vals) per component (or pair counts)."""
        X, Y, Z = dims
        n = X * Y * Z
        boxes = _c(boxes, np.float32).reshape(-1, max(n, 1) if n else 1)
        ncomp = boxes.shape[0]
        comp_ids = np.arange(ncomp, dtype=np.int32) if comp_ids is None else _c(comp_ids, np.int32)
        npairs = np.zeros(ncomp, np.int32)
        if want_pairs:
            runs = np.empty(ncomp * max(n, 1), np.int32)
            vals = np.empty(ncomp * max(n, 1), np.float32)
            rp, vp = runs.ctypes.data, vals.ctypes.data
        else:
            rp = vp = None
        rc = self.lib.wcref_compress(boxes.reshape(-1), ncomp, X, Y, Z, float(keep), t, lev, box_idx,
                                     comp_ids, os.fsencode(out_dir), npairs, rp, vp)
        assert rc == 0
        if not want_pairs:
            return npairs
        return [(runs[c * n:c * n + npairs[c]].copy(), vals[c * n:c * n + npairs[c]].copy())
                for c in range(ncomp)]

    def decompress(self, path, cap=1 << 24):
        out = np.empty(cap, np.float32)
        dims = np.zeros(3, np.int32)
        n = self.lib.wcref_decompress(os.fsencode(path), out, cap, dims)
        assert n >= 0, n
        X, Y, Z = (int(d) for d in dims)
        return out[:n].reshape(Z, Y, X).copy(), (X, Y, Z)

    def rmse(self, actual, pred, dims, ncomp=1):
        X, Y, Z = dims
        out = np.zeros(ncomp, np.float64)
        self.lib.wcref_rmse(_c(actual, np.float32).reshape(-1), _c(pred, np.float32).reshape(-1),
                            ncomp, X, Y, Z, out)
        return out

    def adj_loss(self, rmse, rng):
        return self.lib.wcref_adj_loss(rmse, rng)

    def set_lzma_mode(self, mode):
        self.lib.wcref_set_lzma_mode(int(mode))

    def lzma_seconds(self, reset=False):
        s = self.lib.wcref_lzma_seconds()
        if reset:
            self.lib.wcref_reset_lzma_seconds()
        return s


def have_ref() -> bool:
    return os.path.exists(REF_SO)
