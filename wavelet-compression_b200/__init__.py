"""wavelet-compression_b200 — B200 (sm_100a) numeric core of carsonmw3/wavelet-compression.

    csrc/        hand-written CUDA kernels + the C ABI (include/wcgpu.h) -> libwcgpu.so
    host/        C++ drop-in for the reference's compress()/decompress()/calc_rmse_per_box()
    capi.py      ctypes binding of the C ABI
    core.py      Context / Plan objects
    refapi.py    Python mirror of the reference's three entry points (LZMA + files on the host)
    amr_synth.py synthetic AMReX-shaped workloads
    plotfile.py  AMReX plotfile reader / writer without AMReX (Header, Cell_H, Cell_D)
    sidefiles.py the five .raw side files of a run, byte-compatible with src/readandwrite.cpp
    modes.py     -estimate / -c / -d drivers on top of the GPU path

The directory name contains a hyphen (it mirrors the reference's repository name); import it with
importlib.import_module("wavelet-compression_b200") or through __graft_entry__.package().
"""
from . import amr_synth, capi, modes, plotfile, sidefiles  # noqa: F401

try:  # torch is only needed for the multi-process helpers
    from . import distributed  # noqa: F401
except ImportError:  # pragma: no cover
    distributed = None
from .capi import (WC_DEVICE, WC_F32, WC_F64, WC_HOST, WC_THRESH_GLOBAL,  # noqa: F401
                   WC_THRESH_PER_UNIT, WC_THRESH_QUANTILE, WC_THRESH_QUANTILE_GLOBAL, WcError)
from .core import Context, DecodePlan, PackedUnit, Plan  # noqa: F401
from .refapi import calc_adj_loss, calc_rmse_per_box, compress, decompress  # noqa: F401
