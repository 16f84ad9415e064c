"""Host-side mode drivers on top of the GPU path: Python counterparts of the
reference's `-estimate`, `-c` and `-d` loops (src/modes.cpp:209-328, :24-112, :115-204) that read AMReX
plotfiles with plotfile.py instead of AMReX.  Everything numeric runs on the GPU through the C ABI; LZMA,
file names and directory layout follow the reference (compressed-wavelet-{t}-{level}-{comp}-{box}.xz).
The reference's five .raw side files are replaced by one JSON manifest (metadata, not on the hot path).
"""
from __future__ import annotations

import json
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import plotfile
from .core import Context, PackedUnit
from .refapi import unit_filename, xz_decode, xz_encode

FLT_MAX = float(np.finfo(np.float32).max)
FLT_MIN_POSITIVE = float(np.finfo(np.float32).tiny)   # std::numeric_limits<float>::min(), src/preprocess.cpp:30-31


def component_indices(header: plotfile.Header, components):
    """Exact, case-sensitive name match (src/preprocess.cpp:150-165)."""
    idx = []
    for c in components:
        if c not in header.names:
            raise KeyError(f"component {c!r} not in plotfile Header {header.names}")
        idx.append(header.names.index(c))
    return idx


def fold_minmax(mins, maxs, n_comp):
    """Per-component running min/max over boxes exactly as src/preprocess.cpp:82-88 with its initial values
    (max starts at the smallest POSITIVE float, a reference quirk kept for parity)."""
    lo = [FLT_MAX] * n_comp
    hi = [FLT_MIN_POSITIVE] * n_comp
    for i, (a, b) in enumerate(zip(mins, maxs)):
        c = i % n_comp
        if a < lo[c]:
            lo[c] = float(a)
        if b > hi[c]:
            hi[c] = float(b)
    return lo, hi


def estimate(plt_dir: str, level: int, components, keep: float, ctx: Context | None = None, threads: int = 0):
    """`-estimate` (src/modes.cpp:209-328): ONE file, ONE level, all requested components.
    Returns {component: {rmse, adjusted_loss}}, compressed_percent and the per-unit pair counts."""
    ctx = ctx or Context(0)
    hdr = plotfile.read_header(plt_dir)
    comp_idxs = component_indices(hdr, components)
    lev = plotfile.read_level(plt_dir, level)
    units = plotfile.level_units(lev, comp_idxs)
    boxes = [u[0] for u in units]
    dims = [u[1] for u in units]
    nc = len(comp_idxs)
    packed = ctx.compress_batch(boxes, keep, dims=dims)                  # F, T, M, P on the GPU (float64 ingest)
    recon = ctx.decompress_batch(packed)                                 # U, I
    rmse = ctx.rmse_batch(boxes, recon)                                  # R (narrowing of `actual` on the GPU)
    mins, maxs = ctx.minmax_batch(boxes, dims)
    lo, hi = fold_minmax(mins, maxs, nc)
    with ThreadPoolExecutor(threads or os.cpu_count() or 1) as pool:     # LZMA stays on the host
        xz_sizes = list(pool.map(lambda p: len(xz_encode(p.serialize())), packed))
    out = {"components": {}, "npairs": [p.npairs for p in packed]}
    for c, name in enumerate(components):
        per_box = rmse[c::nc]
        mean_rmse = float(np.sum(per_box) / len(per_box))                # std::accumulate / size, src/modes.cpp:284-285
        out["components"][name] = {"rmse": mean_rmse, "adjusted_loss": mean_rmse / (hi[c] - lo[c]) if hi[c] != lo[c] else
                                   float(np.float64(mean_rmse) / np.float64(hi[c] - lo[c]) if mean_rmse else np.nan),
                                   "min": lo[c], "max": hi[c]}
    ldir = os.path.join(plt_dir, f"Level_{level}")
    raw_size = float(sum(os.path.getsize(os.path.join(ldir, f)) for f in os.listdir(ldir)))
    raw_size = raw_size / len(hdr.names) * nc                             # src/modes.cpp:316-318
    out["compressed_percent"] = float(sum(xz_sizes)) / raw_size * 100.0   # src/modes.cpp:321-324
    return out


def compress_run(plt_dirs, levels, components, keep: float, compressed_dir: str, ctx: Context | None = None,
                 threads: int = 0):
    """`-c` (src/modes.cpp:24-112): every (file, level, box, component) in one GPU batch, one .xz per unit."""
    ctx = ctx or Context(0)
    os.makedirs(compressed_dir, exist_ok=True)
    keys, boxes, dims = [], [], []
    manifest = {"files": [os.path.basename(os.path.normpath(p)) for p in plt_dirs], "levels": list(levels),
                "components": list(components), "boxes": {}}
    hdr0 = None
    for t, plt in enumerate(plt_dirs):
        hdr = plotfile.read_header(plt)
        hdr0 = hdr0 or hdr
        comp_idxs = component_indices(hdr, components)
        for li, level in enumerate(levels):
            lev = plotfile.read_level(plt, level)
            manifest["boxes"][f"{t}-{li}"] = [[list(f.lo), list(f.hi)] for f in lev.fabs]
            for b, fab in enumerate(lev.fabs):
                for c in comp_idxs:
                    keys.append((t, li, c, b))
                    boxes.append(fab.data[c])
                    dims.append(fab.dims)
    manifest["comp_idxs"] = component_indices(hdr0, components)
    packed = ctx.compress_batch(boxes, keep, dims=dims)

    def write(i):
        t, li, c, b = keys[i]
        with open(os.path.join(compressed_dir, unit_filename(t, li, c, b)), "wb") as f:
            f.write(xz_encode(packed[i].serialize()))
    with ThreadPoolExecutor(threads or os.cpu_count() or 1) as pool:
        list(pool.map(write, range(len(keys))))
    with open(os.path.join(compressed_dir, "wcgpu_manifest.json"), "w") as f:
        json.dump(manifest, f)
    return manifest


def decompress_run(compressed_dir: str, out_dir: str, ctx: Context | None = None, threads: int = 0):
    """`-d` (src/modes.cpp:115-204): reads every unit file, one GPU batch for U + I, writes Level_k/Cell_*
    for each timestep (float32 results widened to float64 on the GPU, as src/writeplotfile.cpp:103)."""
    ctx = ctx or Context(0)
    m = json.load(open(os.path.join(compressed_dir, "wcgpu_manifest.json")))
    keys = []
    for key, bl in m["boxes"].items():
        t, li = (int(v) for v in key.split("-"))
        for b in range(len(bl)):
            for c in m["comp_idxs"]:
                keys.append((t, li, c, b))

    def read(k):
        with open(os.path.join(compressed_dir, unit_filename(*k)), "rb") as f:
            return PackedUnit.deserialize(xz_decode(f.read()))
    with ThreadPoolExecutor(threads or os.cpu_count() or 1) as pool:
        packed = list(pool.map(read, keys))
    recon = ctx.decompress_batch(packed, out_dtype=np.float64)
    nc = len(m["comp_idxs"])
    by = {}
    for k, r in zip(keys, recon):
        by.setdefault((k[0], k[1]), {}).setdefault(k[3], []).append(r)
    for (t, li), boxes in by.items():
        bl = m["boxes"][f"{t}-{li}"]
        data = [np.stack(boxes[b]) for b in range(len(bl))]
        plotfile.write_level(os.path.join(out_dir, m["files"][t]), m["levels"][li],
                             [(tuple(lo), tuple(hi)) for lo, hi in bl], data, nc)
    return m
