"""Host-side objects over the C ABI: Context (one per GPU) and Plan (one per batch of units).

Boxes follow the reference's Grid3D layout (src/grid.h:18): x fastest.  A NumPy / torch box is
therefore indexed [k][j][i] and has shape (nz, ny, nx); `dims` are always written (nx, ny, nz).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import (BOX_DESC, PACKED, PAIR, WC_DEVICE, WC_F32, WC_F64, WC_HOST, WC_THRESH_GLOBAL,
                   WC_THRESH_PER_UNIT, check)


@dataclass
class PackedUnit:
    """One unit's (run, value) pairs on the host: the reference's CompressedWavelet
    (src/box-structs.h:65-70)."""
    dims: tuple
    ncoef: int
    runs: np.ndarray  # int32[K]
    vals: np.ndarray  # float32[K]

    @property
    def npairs(self) -> int:
        return int(self.runs.size)

    def serialize(self) -> bytes:
        """The 20+8K-byte buffer of serialize_compressed_wavelet (src/compressor.cpp:55-80)."""
        head = np.array([*self.dims, self.ncoef, self.npairs], dtype="<i4").tobytes()
        body = np.empty(self.npairs, PAIR)
        body["run"] = self.runs
        body["val"] = self.vals
        return head + body.tobytes()

    @staticmethod
    def deserialize(buf: bytes) -> "PackedUnit":
        """deserialize_compressed_wavelet (src/decompressor.cpp:35-74)."""
        h = np.frombuffer(buf, "<i4", 5)
        k = int(h[4])
        body = np.frombuffer(buf, PAIR, k, 20)
        return PackedUnit((int(h[0]), int(h[1]), int(h[2])), int(h[3]), body["run"].copy(),
                          body["val"].copy())


def _np_dtype_code(a: np.ndarray) -> int:
    if a.dtype == np.float32:
        return WC_F32
    if a.dtype == np.float64:
        return WC_F64
    raise TypeError(f"boxes must be float32 or float64, got {a.dtype}")


def _host_descs(boxes, dims=None):
    """wc_box_desc table for a list of C-contiguous host arrays shaped (nz, ny, nx)."""
    keep = []
    ptrs, dts, dd = [], [], []
    for i, b in enumerate(boxes):
        b = np.ascontiguousarray(b)
        keep.append(b)
        if dims is not None:
            d = tuple(dims[i])
        else:
            if b.ndim != 3:
                raise ValueError("box arrays must be 3-D (nz, ny, nx) unless dims are given")
            d = (b.shape[2], b.shape[1], b.shape[0])
        if b.size != d[0] * d[1] * d[2]:
            raise ValueError("box size does not match dims")
        ptrs.append(b.ctypes.data if b.size else 0)
        dts.append(_np_dtype_code(b))
        dd.append(d)
    return capi.box_descs(ptrs, dts, dd) if boxes else np.zeros(0, BOX_DESC), keep


class Context:
    """wc_ctx: one per GPU.  `stream` is an optional cudaStream_t address (e.g.
    torch.cuda.current_stream().cuda_stream) on which all work is then issued."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = capi.load()
        h = C.c_void_p()
        if stream is None:
            check(self.lib.wc_create(C.byref(h), device), "wc_create")
        else:
            check(self.lib.wc_create_on_stream(C.byref(h), device, C.c_void_p(stream)),
                  "wc_create_on_stream")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.wc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- misc ---------------------------------------------------------------------------------
    def sync(self):
        check(self.lib.wc_sync(self.h), "wc_sync", self.h)

    def set_path(self, path: int):
        """0 auto, 1 generic only, 2 fused only."""
        check(self.lib.wc_set_option(self.h, capi.WC_OPT_PATH, path), "wc_set_option", self.h)

    def counter(self, which: int) -> int:
        v = C.c_uint64(0)
        check(self.lib.wc_get_counter(self.h, which, C.byref(v)), "wc_get_counter", self.h)
        return v.value

    def set_option(self, option: int, value: int):
        check(self.lib.wc_set_option(self.h, option, int(value)), "wc_set_option", self.h)

    def set_profile(self, on: bool):
        check(self.lib.wc_set_option(self.h, capi.WC_OPT_PROFILE, 1 if on else 0), "wc_set_option", self.h)

    def kernel_stats(self) -> dict:
        """{kernel name: (launches, total device ms)} since the last reset_counters()."""
        out, i = {}, 0
        while True:
            name, ms, n = C.c_char_p(), C.c_double(0), C.c_uint64(0)
            if self.lib.wc_kernel_stats(self.h, i, C.byref(name), C.byref(ms), C.byref(n)) != 0:
                break
            if n.value:
                out[name.value.decode()] = (n.value, ms.value)
            i += 1
        return out

    def reset_counters(self):
        check(self.lib.wc_reset_counters(self.h), "wc_reset_counters", self.h)

    # -- blocking batch API (host arrays) ------------------------------------------------------
    def compress_batch(self, boxes, keep: float, thresh_mode: int = WC_THRESH_PER_UNIT,
                       dims=None) -> list[PackedUnit]:
        """Numeric part of compress() (src/compressor.cpp:203-247) for a list of host boxes."""
        descs, hold = _host_descs(boxes, dims)
        n = len(descs)
        out = np.zeros(max(n, 1), PACKED)
        check(self.lib.wc_compress_batch(self.h, descs.ctypes.data, n, WC_HOST, float(keep),
                                         thresh_mode, out.ctypes.data, WC_HOST),
              "wc_compress_batch", self.h)
        return [self._packed_to_host(out[i]) for i in range(n)]

    @staticmethod
    def _packed_to_host(rec) -> PackedUnit:
        k = int(rec["npairs"])
        if k:
            buf = (C.c_char * (8 * k)).from_address(int(rec["pairs"]))
            pr = np.frombuffer(buf, PAIR, k)
            runs, vals = pr["run"].copy(), pr["val"].copy()
        else:
            runs, vals = np.zeros(0, np.int32), np.zeros(0, np.float32)
        return PackedUnit(tuple(int(s) for s in rec["shape"]), int(rec["ncoef"]), runs, vals)

    def decompress_batch(self, packed: list[PackedUnit], out_dtype=np.float32) -> list[np.ndarray]:
        """Numeric part of decompress() (src/decompressor.cpp:245-254)."""
        n = len(packed)
        rec = np.zeros(max(n, 1), PACKED)
        hold, outs = [], []
        for i, p in enumerate(packed):
            pr = np.empty(p.npairs, PAIR)
            pr["run"], pr["val"] = p.runs, p.vals
            hold.append(pr)
            rec[i]["shape"] = p.dims
            rec[i]["ncoef"] = p.ncoef
            rec[i]["npairs"] = p.npairs
            rec[i]["pairs"] = pr.ctypes.data if p.npairs else 0
            X, Y, Z = p.dims
            outs.append(np.empty((Z, Y, X), out_dtype))
        od, _ = _host_descs(outs, [p.dims for p in packed])
        check(self.lib.wc_decompress_batch(self.h, rec.ctypes.data, n, WC_HOST, od.ctypes.data,
                                           WC_HOST), "wc_decompress_batch", self.h)
        return outs

    def rmse_batch(self, actual, pred) -> np.ndarray:
        """calc_rmse_per_box (src/calc-loss.cpp:12-43), one value per (actual, pred) pair."""
        da, ha = _host_descs(actual)
        db, hb = _host_descs(pred)
        n = len(da)
        out = np.zeros(max(n, 1), np.float64)
        check(self.lib.wc_rmse_batch(self.h, da.ctypes.data, db.ctypes.data, n, WC_HOST,
                                     out.ctypes.data), "wc_rmse_batch", self.h)
        return out[:n]

    def minmax_batch(self, boxes, dims=None):
        """Per-unit (min, max) of the narrowed float32 values (src/preprocess.cpp:82-88)."""
        d, hold = _host_descs(boxes, dims)
        n = len(d)
        lo, hi = np.zeros(max(n, 1), np.float32), np.zeros(max(n, 1), np.float32)
        check(self.lib.wc_minmax_batch(self.h, d.ctypes.data, n, WC_HOST, lo.ctypes.data, hi.ctypes.data),
              "wc_minmax_batch", self.h)
        return lo[:n], hi[:n]

    # -- un-fused primitives -------------------------------------------------------------------
    def haar_forward(self, box: np.ndarray) -> np.ndarray:
        d, hold = _host_descs([box])
        out = np.empty(max(hold[0].size, 1), np.float32)
        check(self.lib.wc_haar_forward(self.h, d.ctypes.data, WC_HOST, out.ctypes.data),
              "wc_haar_forward", self.h)
        return out[:hold[0].size]

    def haar_inverse(self, flat: np.ndarray, dims) -> np.ndarray:
        X, Y, Z = dims
        flat = np.ascontiguousarray(flat, np.float32)
        out = np.empty(max(X * Y * Z, 1), np.float32)
        check(self.lib.wc_haar_inverse(self.h, flat.ctypes.data, X, Y, Z, WC_HOST, out.ctypes.data),
              "wc_haar_inverse", self.h)
        return out[:X * Y * Z].reshape(Z, Y, X)

    def threshold_pack(self, flat: np.ndarray, keep: float):
        flat = np.ascontiguousarray(flat, np.float32)
        pr = np.empty(max(flat.size, 1), PAIR)
        k = C.c_int32(0)
        check(self.lib.wc_threshold_pack(self.h, flat.ctypes.data, flat.size, float(keep), WC_HOST,
                                         pr.ctypes.data, C.byref(k)), "wc_threshold_pack", self.h)
        return pr["run"][:k.value].copy(), pr["val"][:k.value].copy()

    def rle_decode(self, runs, vals, total: int) -> np.ndarray:
        pr = np.empty(max(len(runs), 1), PAIR)
        pr["run"][:len(runs)] = runs
        pr["val"][:len(runs)] = vals
        out = np.empty(max(total, 1), np.float32)
        check(self.lib.wc_rle_decode(self.h, pr.ctypes.data, len(runs), total, WC_HOST,
                                     out.ctypes.data), "wc_rle_decode", self.h)
        return out[:total]

    # -- plans ----------------------------------------------------------------------------------
    def plan(self, descs: np.ndarray, in_space: int) -> "Plan":
        return Plan(self, descs, in_space)

    def decode_plan(self, out_descs: np.ndarray, out_space: int) -> "DecodePlan":
        return DecodePlan(self, out_descs, out_space)

    def plan_host(self, boxes, dims=None) -> "Plan":
        descs, hold = _host_descs(boxes, dims)
        p = Plan(self, descs, WC_HOST)
        p._hold = hold
        return p


class Plan:
    """wc_plan: fixed unit list, device-resident buffers, asynchronous execution."""

    def __init__(self, ctx: Context, descs: np.ndarray, in_space: int):
        assert descs.dtype == BOX_DESC
        self.ctx, self.lib = ctx, ctx.lib
        self.n = len(descs)
        self.descs = descs.copy()
        self.in_space = in_space
        h = C.c_void_p()
        check(self.lib.wc_plan_create(ctx.h, self.descs.ctypes.data, self.n, in_space, C.byref(h)),
              "wc_plan_create", ctx.h)
        self.h = h
        self._hold = None
        self._packed = np.zeros(max(self.n, 1), PACKED)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.lib.wc_plan_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_inputs(self, descs: np.ndarray):
        assert len(descs) == self.n
        self.descs = descs.copy()
        check(self.lib.wc_plan_set_inputs(self.h, self.descs.ctypes.data), "wc_plan_set_inputs",
              self.ctx.h)

    def compress(self, keep: float, thresh_mode: int = WC_THRESH_PER_UNIT):
        check(self.lib.wc_plan_compress(self.h, float(keep), thresh_mode), "wc_plan_compress",
              self.ctx.h)

    def transform(self) -> int:
        """Forward transform + batch arg-max key; returns the device address of the uint64 key."""
        k = C.c_void_p()
        check(self.lib.wc_plan_transform(self.h, C.byref(k)), "wc_plan_transform", self.ctx.h)
        return k.value

    def pack_with_key(self, keep: float, key_dev: int):
        check(self.lib.wc_plan_pack_with_key(self.h, float(keep), C.c_void_p(key_dev)),
              "wc_plan_pack_with_key", self.ctx.h)

    # EXTENSION: quantile thresholds (radix select), split at the histograms for multi-GPU runs
    def quantile_begin(self, keep: float, global_mode: bool, n_total: int = 0):
        check(self.lib.wc_plan_quantile_begin(self.h, float(keep), int(global_mode), int(n_total)),
              "wc_plan_quantile_begin", self.ctx.h)

    def quantile_hist(self, pass_: int) -> int:
        """Histogram of radix-select pass 0, 1 or 2; returns the device address of the uint64[2048] row(s)."""
        h = C.c_void_p()
        check(self.lib.wc_plan_quantile_hist(self.h, int(pass_), C.byref(h)), "wc_plan_quantile_hist", self.ctx.h)
        return h.value

    def quantile_pick(self, pass_: int):
        check(self.lib.wc_plan_quantile_pick(self.h, int(pass_)), "wc_plan_quantile_pick", self.ctx.h)

    def quantile_pack(self):
        check(self.lib.wc_plan_quantile_pack(self.h), "wc_plan_quantile_pack", self.ctx.h)

    def total_pairs(self) -> int:
        t = C.c_int64(0)
        check(self.lib.wc_plan_total_pairs(self.h, C.byref(t)), "wc_plan_total_pairs", self.ctx.h)
        return t.value

    def fetch_records(self, space: int) -> np.ndarray:
        """wc_packed records; .pairs are device (slot) or pinned-host (dense) addresses."""
        check(self.lib.wc_plan_fetch(self.h, self._packed.ctypes.data, space), "wc_plan_fetch",
              self.ctx.h)
        return self._packed[:self.n]

    def compress_to_host_records(self, keep: float) -> np.ndarray:
        """Pipelined H2D / kernels / D2H (host-resident inputs): wc_packed records with pinned-host pairs."""
        check(self.lib.wc_plan_compress_to_host(self.h, float(keep), self._packed.ctypes.data),
              "wc_plan_compress_to_host", self.ctx.h)
        return self._packed[:self.n]

    def compress_to_host_chunked(self, keep: float, on_chunk) -> np.ndarray:
        """wc_plan_compress_to_host_chunked: on_chunk(first_unit, n_units, records) is called on this thread as soon
        as the pairs of those units are in pinned host memory, while later chunks are still on the GPU."""
        recs = self._packed

        def tramp(_user, first, n, _units):
            on_chunk(first, n, recs[first:first + n])
        cb = capi.CHUNK_FN(tramp)
        check(self.lib.wc_plan_compress_to_host_chunked(self.h, float(keep), self._packed.ctypes.data, cb, None),
              "wc_plan_compress_to_host_chunked", self.ctx.h)
        return self._packed[:self.n]

    def unit_stats(self, minmax: bool = True):
        """(mins, maxs, need32) of the last compress; mins / maxs need WC_OPT_INGEST_STATS at compress time."""
        n = max(self.n, 1)
        lo, hi, n32 = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        check(self.lib.wc_plan_unit_stats(self.h, lo.ctypes.data if minmax else None, hi.ctypes.data if minmax else None,
                                          n32.ctypes.data), "wc_plan_unit_stats", self.ctx.h)
        return lo[:self.n], hi[:self.n], n32[:self.n]

    def fetch_host(self) -> list[PackedUnit]:
        rec = self.fetch_records(WC_HOST)
        return [Context._packed_to_host(rec[i]) for i in range(self.n)]

    def decompress(self, out_descs: np.ndarray, out_space: int):
        assert len(out_descs) == self.n
        check(self.lib.wc_plan_decompress(self.h, out_descs.ctypes.data, out_space),
              "wc_plan_decompress", self.ctx.h)

    def rmse(self, recon_descs: np.ndarray) -> np.ndarray:
        out = np.zeros(max(self.n, 1), np.float64)
        check(self.lib.wc_plan_rmse(self.h, recon_descs.ctypes.data, out.ctypes.data),
              "wc_plan_rmse", self.ctx.h)
        return out[:self.n]


class DecodePlan:
    """wc_dplan: fixed output boxes; every decode() takes the batch as one dense pair stream + per-unit counts."""

    def __init__(self, ctx: Context, out_descs: np.ndarray, out_space: int):
        assert out_descs.dtype == BOX_DESC
        self.ctx, self.lib = ctx, ctx.lib
        self.n = len(out_descs)
        self.descs = out_descs.copy()
        h = C.c_void_p()
        check(self.lib.wc_dplan_create(ctx.h, self.descs.ctypes.data, self.n, out_space, C.byref(h)),
              "wc_dplan_create", ctx.h)
        self.h = h

    def decode(self, pairs_addr: int, npairs_addr: int, in_space: int):
        check(self.lib.wc_dplan_decode(self.h, C.c_void_p(pairs_addr), C.c_void_p(npairs_addr), in_space),
              "wc_dplan_decode", self.ctx.h)

    def finish(self):
        check(self.lib.wc_dplan_finish(self.h), "wc_dplan_finish", self.ctx.h)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.lib.wc_dplan_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
