"""ctypes binding of libwcgpu.so (include/wcgpu.h) — the only way this package computes anything.

There is no Python/NumPy/torch implementation of the numeric core in this package and no CPU
fallback: if libwcgpu.so is missing or no B200 is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# WCGPU_LIB: developer hook for instrumented builds of the SAME library (make PHASE_PROFILE=1 OUT=...); it names
# another libwcgpu, never another implementation
LIB_PATH = os.environ.get("WCGPU_LIB") or os.path.join(PKG_DIR, "libwcgpu.so")

WC_OK = 0
WC_F32, WC_F64 = 0, 1
WC_HOST, WC_DEVICE = 0, 1
WC_THRESH_PER_UNIT, WC_THRESH_GLOBAL, WC_THRESH_QUANTILE, WC_THRESH_QUANTILE_GLOBAL = 0, 1, 2, 3
WC_OPT_PATH, WC_OPT_PROFILE, WC_OPT_OVERLAP, WC_OPT_SEG_INDEX, WC_OPT_COPY_ONLY, WC_OPT_INGEST_STATS = 0, 1, 2, 3, 4, 5
WC_OPT_DECODE_PIPE = 6
WC_PACKED_NEED32 = 1
WC_CTR_KERNEL_LAUNCHES, WC_CTR_H2D_BYTES, WC_CTR_D2H_BYTES = 0, 1, 2

# numpy mirrors of the POD structs (layout checked against sizeof in tests/test_abi.py)
BOX_DESC = np.dtype([("data", "<u8"), ("dtype", "<i4"), ("nx", "<i4"), ("ny", "<i4"), ("nz", "<i4")],
                    align=True)
PACKED = np.dtype([("shape", "<i4", (3,)), ("ncoef", "<i4"), ("npairs", "<i4"), ("flags", "<i4"),
                   ("pairs", "<u8")], align=True)
PAIR = np.dtype([("run", "<i4"), ("val", "<f4")])
assert BOX_DESC.itemsize == 24 and PACKED.itemsize == 32 and PAIR.itemsize == 8

# every symbol include/wcgpu.h declares: (restype, argtypes)
_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
SIGNATURES = {
    "wc_version": (_i, []),
    "wc_strerror": (C.c_char_p, [_i]),
    "wc_box_kernel_class": (C.c_char_p, [_i, _i, _i, _i, _i]),
    "wc_device_count": (_i, [C.POINTER(_i)]),
    "wc_create": (_i, [C.POINTER(_vp), _i]),
    "wc_create_on_stream": (_i, [C.POINTER(_vp), _i, _vp]),
    "wc_destroy": (_i, [_vp]),
    "wc_sync": (_i, [_vp]),
    "wc_last_error": (C.c_char_p, [_vp]),
    "wc_set_option": (_i, [_vp, _i, C.c_int64]),
    "wc_get_counter": (_i, [_vp, _i, C.POINTER(C.c_uint64)]),
    "wc_reset_counters": (_i, [_vp]),
    "wc_kernel_stats": (_i, [_vp, _i, C.POINTER(C.c_char_p), C.POINTER(_d), C.POINTER(C.c_uint64)]),
    "wc_host_alloc": (_i, [C.POINTER(_vp), C.c_size_t]),
    "wc_host_free": (_i, [_vp]),
    "wc_device_alloc": (_i, [_vp, C.POINTER(_vp), C.c_size_t]),
    "wc_device_free": (_i, [_vp, _vp]),
    "wc_memcpy": (_i, [_vp, _vp, _vp, C.c_size_t, _i]),
    "wc_compress_batch": (_i, [_vp, _vp, _i, _i, _d, _i, _vp, _i]),
    "wc_decompress_batch": (_i, [_vp, _vp, _i, _i, _vp, _i]),
    "wc_rmse_batch": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "wc_minmax_batch": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "wc_haar_forward": (_i, [_vp, _vp, _i, _vp]),
    "wc_haar_inverse": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "wc_threshold_pack": (_i, [_vp, _vp, _i, _d, _i, _vp, C.POINTER(C.c_int32)]),
    "wc_rle_decode": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "wc_serialize_header": (_i, [_vp, _vp]),
    "wc_plan_create": (_i, [_vp, _vp, _i, _i, C.POINTER(_vp)]),
    "wc_plan_destroy": (_i, [_vp]),
    "wc_plan_set_inputs": (_i, [_vp, _vp]),
    "wc_plan_compress": (_i, [_vp, _d, _i]),
    "wc_plan_fetch": (_i, [_vp, _vp, _i]),
    "wc_plan_compress_to_host": (_i, [_vp, _d, _vp]),
    "wc_plan_compress_to_host_chunked": (_i, [_vp, _d, _vp, _vp, _vp]),
    "wc_plan_unit_stats": (_i, [_vp, _vp, _vp, _vp]),
    "wc_dplan_create": (_i, [_vp, _vp, _i, _i, C.POINTER(_vp)]),
    "wc_dplan_destroy": (_i, [_vp]),
    "wc_dplan_decode": (_i, [_vp, _vp, _vp, _i]),
    "wc_dplan_finish": (_i, [_vp]),
    "wc_plan_total_pairs": (_i, [_vp, C.POINTER(C.c_int64)]),
    "wc_plan_decompress": (_i, [_vp, _vp, _i]),
    "wc_plan_rmse": (_i, [_vp, _vp, _vp]),
    "wc_plan_transform": (_i, [_vp, C.POINTER(_vp)]),
    "wc_plan_pack_with_key": (_i, [_vp, _d, _vp]),
    "wc_plan_quantile_begin": (_i, [_vp, _d, _i, C.c_uint64]),
    "wc_plan_quantile_hist": (_i, [_vp, _i, C.POINTER(_vp)]),
    "wc_plan_quantile_pick": (_i, [_vp, _i]),
    "wc_plan_quantile_pack": (_i, [_vp]),
}


CHUNK_FN = C.CFUNCTYPE(None, _vp, _i, _i, _vp)   # wc_chunk_fn


class WcError(RuntimeError):
    def __init__(self, status: int, what: str, detail: str = ""):
        self.status = status
        super().__init__(f"{what}: wc_status {status}" + (f" ({detail})" if detail else ""))


_lib = None


def load(path: str | None = None) -> C.CDLL:
    """dlopen libwcgpu.so and bind every declared symbol.  Raises if the library was not built —
    run `python -c "import __graft_entry__ as g; g.build()"` or `make -C wavelet-compression_b200/csrc`."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            f"{p} not found: the CUDA extension is not built and this package has no fallback path")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def strerror(status: int) -> str:
    return load().wc_strerror(status).decode()


def check(status: int, what: str, ctx=None):
    if status != WC_OK:
        detail = strerror(status)
        if ctx:
            le = load().wc_last_error(ctx).decode()
            if le:
                detail += "; " + le
        raise WcError(status, what, detail)


def box_descs(ptrs, dtypes, dims) -> np.ndarray:
    """Array of wc_box_desc / wc_box_out from parallel sequences (addresses, wc_dtype, (nx,ny,nz))."""
    n = len(ptrs)
    a = np.zeros(n, BOX_DESC)
    a["data"] = np.asarray(ptrs, dtype=np.uint64)
    a["dtype"] = np.asarray(dtypes, dtype=np.int32)
    d = np.asarray(dims, dtype=np.int32).reshape(n, 3)
    a["nx"], a["ny"], a["nz"] = d[:, 0], d[:, 1], d[:, 2]
    return a
