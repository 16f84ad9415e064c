"""Python mirror of the reference's three hot-path entry points, on top of the C ABI.

Same names, argument order and meaning as the C++ functions src/modes.cpp calls, so the parity
tests read like the reference's own doctests (src/compressor.cpp:387-406, src/calc-loss.cpp:68-86):

    compress(box, components, keep, time, level, box_index, compressed_dir)   src/compressor.h:9-15
    decompress(file_path, time, level, component, box_idx) -> box             src/decompressor.h:6-10
    calc_rmse_per_box(actual, pred, num_components) -> [rmse]                  src/calc-loss.h:6-8

The numeric core runs on the GPU (libwcgpu); what stays on the host is exactly what stays on the
host in the reference: the 20-byte header, the .xz container (xz preset 6, CRC64, one stream —
src/compressor.cpp:260-285; Python's lzma module wraps the same liblzma) and the file name
compressed-wavelet-{t}-{level}-{component}-{box}.xz (src/compressor.cpp:250-254).
"""
from __future__ import annotations

import lzma
import os

import numpy as np

from .core import Context, PackedUnit

_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None or _default_ctx.h is None:
        _default_ctx = Context(0)
    return _default_ctx


def xz_encode(buf: bytes) -> bytes:
    """lzma_easy_encoder(preset 6, LZMA_CHECK_CRC64) + lzma_code(FINISH), src/compressor.cpp:260-285."""
    return lzma.compress(buf, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC64, preset=6)


def xz_decode(buf: bytes) -> bytes:
    """lzma_stream_decoder(LZMA_CONCATENATED), src/decompressor.cpp:188-220."""
    return lzma.decompress(buf, format=lzma.FORMAT_XZ)


def unit_filename(time: int, level: int, component: int, box_index: int) -> str:
    return f"compressed-wavelet-{time}-{level}-{component}-{box_index}.xz"


def compress(box, components, keep, time, level, box_index, compressed_dir, ctx: Context | None = None):
    """box: the multiBox3D — a sequence of per-component arrays shaped (nz, ny, nx), float32 (or the
    raw float64 FAB slabs).  components: the Header indices used in the file names.  Writes one .xz
    per component and returns the list of PackedUnit (the reference returns CompressedWavelet)."""
    ctx = ctx or default_context()
    boxes = [np.asarray(box[c]) for c in range(len(components))]
    packed = ctx.compress_batch(boxes, float(keep))
    for c, p in zip(components, packed):
        path = os.path.join(compressed_dir, unit_filename(time, level, c, box_index))
        with open(path, "wb") as f:
            f.write(xz_encode(p.serialize()))
    return packed


def decompress(file_path, time=0, level=0, component=0, box_idx=0, ctx: Context | None = None):
    """Returns the regenerated box, float32 (nz, ny, nx).  The four ints are unused, as in the
    reference (src/decompressor.cpp:238-255)."""
    ctx = ctx or default_context()
    with open(file_path, "rb") as f:
        p = PackedUnit.deserialize(xz_decode(f.read()))
    return ctx.decompress_batch([p])[0]


def calc_rmse_per_box(actual, pred, num_components, ctx: Context | None = None):
    ctx = ctx or default_context()
    a = [np.asarray(actual[c], np.float32) for c in range(num_components)]
    b = [np.asarray(pred[c], np.float32) for c in range(num_components)]
    return list(ctx.rmse_batch(a, b))


def calc_adj_loss(rmse: float, value_range: float) -> float:
    """src/calc-loss.cpp:49-51 (one host divide; numpy semantics so a zero range gives inf/nan like C++)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.float64(rmse) / np.float64(value_range))
