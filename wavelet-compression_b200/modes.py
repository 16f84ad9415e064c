"""Host-side mode drivers on top of the GPU path: Python counterparts of the reference's
`-estimate`, `-c` and `-d` loops (src/modes.cpp:209-328, :24-112, :115-204) that read AMReX plotfiles with
plotfile.py instead of AMReX.  Everything numeric runs on the GPU through the C ABI; what stays on the host is what
the reference keeps there: LZMA, the file names (compressed-wavelet-{t}-{level}-{comp}-{box}.xz), the five .raw side
files (sidefiles.py, byte-compatible with src/readandwrite.cpp) and the plotfile directory layout.

  estimate       one H2D of the raw float64 FAB slabs; compress -> decompress -> RMSE -> min/max all stay in HBM
                 (wc_plan_*); only the packed pairs come back, chunk by chunk, for the LZMA size estimate, which runs
                 on host threads WHILE the GPU works on later chunks.
  compress_run   same pipeline, the chunk callback feeds the .xz writers.
  decompress_run the files' pair bytes are concatenated into one dense stream and decoded by a decode plan
                 (wc_dplan_*), float64 out, then written as a complete plotfile (Header + Level_k/Cell_H + Cell_D).
"""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import capi, plotfile, sidefiles
from .core import Context, _host_descs
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


def sequential_mean(values) -> float:
    """std::accumulate(begin, end, 0.0) / size (src/modes.cpp:284-285): a left-to-right float64 sum."""
    s = 0.0
    for v in values:
        s += float(v)
    return s / len(values)


def _serialized(rec) -> bytes:
    """serialize_compressed_wavelet (src/compressor.cpp:55-80) of one wc_packed record with host pairs."""
    k = int(rec["npairs"])
    head = np.array([*[int(s) for s in rec["shape"]], int(rec["ncoef"]), k], dtype="<i4").tobytes()
    return head + (C.string_at(int(rec["pairs"]), 8 * k) if k else b"")


def _compress_units(ctx: Context, boxes, dims, keep: float, per_unit, threads: int, ingest_stats: bool):
    """One plan over host boxes: pipelined H2D / kernels / D2H; `per_unit(i, serialized_bytes)` runs on a host thread
    pool fed by the chunk callback, i.e. while later chunks are still on the GPU.  Returns the plan (its packed result
    stays in HBM for the round trip of estimate)."""
    descs, hold = _host_descs(boxes, dims)
    if ingest_stats:
        ctx.set_option(capi.WC_OPT_INGEST_STATS, 1)
    try:
        plan = ctx.plan(descs, capi.WC_HOST)
        plan._hold = hold
        with ThreadPoolExecutor(threads or os.cpu_count() or 1) as pool:
            futs = []

            def on_chunk(first, n, recs):
                for j in range(n):
                    futs.append(pool.submit(per_unit, first + j, _serialized(recs[j])))
            rec = plan.compress_to_host_chunked(keep, on_chunk).copy()
            for f in futs:
                f.result()
    finally:
        if ingest_stats:
            ctx.set_option(capi.WC_OPT_INGEST_STATS, 0)
    return plan, rec


def estimate(plt_dir: str, level: int, components, keep: float, ctx: Context | None = None, threads: int = 0):
    """`-estimate` (src/modes.cpp:209-328): ONE file, ONE level, all requested components.
    Returns {component: {rmse, adjusted_loss}}, compressed_percent, the per-unit pair counts and need32 flags."""
    ctx = ctx or Context(0)
    hdr = plotfile.read_header(plt_dir)
    comp_idxs = component_indices(hdr, components)
    lev = plotfile.read_level(plt_dir, level)
    units = plotfile.level_units(lev, comp_idxs)
    boxes = [u[0] for u in units]
    dims = [u[1] for u in units]
    nc, n = len(comp_idxs), len(units)
    xz_sizes = [0] * n

    def size_of(i, ser):
        xz_sizes[i] = len(xz_encode(ser))                                  # LZMA stays on the host (threads)
    h2d0 = ctx.counter(capi.WC_CTR_H2D_BYTES)
    plan, rec = _compress_units(ctx, boxes, dims, keep, size_of, threads, ingest_stats=True)   # F, T, M, P (float64 ingest)
    # U, I, R without leaving HBM: reconstruction into device boxes, RMSE against the plan's own inputs
    ncoef = [d[0] * d[1] * d[2] for d in dims]
    dbuf = C.c_void_p()
    capi.check(ctx.lib.wc_device_alloc(ctx.h, C.byref(dbuf), 4 * max(sum(ncoef), 1)), "wc_device_alloc", ctx.h)
    try:
        offs = np.concatenate([[0], np.cumsum(ncoef)])[:-1]
        od = capi.box_descs([dbuf.value + 4 * int(o) for o in offs], [capi.WC_F32] * n, dims)
        plan.decompress(od, capi.WC_DEVICE)
        rmse = plan.rmse(od)
        mins, maxs, need32 = plan.unit_stats()
    finally:
        ctx.lib.wc_device_free(ctx.h, dbuf)
    h2d = ctx.counter(capi.WC_CTR_H2D_BYTES) - h2d0
    plan.close()
    lo, hi = fold_minmax(mins, maxs, nc)
    out = {"components": {}, "npairs": [int(k) for k in rec["npairs"]], "need32": [bool(x) for x in need32],
           "h2d_bytes": int(h2d), "input_bytes": int(sum(b.nbytes for b in boxes))}
    for c, name in enumerate(components):
        mean_rmse = sequential_mean(rmse[c::nc])                          # src/modes.cpp:284-285
        with np.errstate(divide="ignore", invalid="ignore"):
            adj = float(np.float64(mean_rmse) / np.float64(np.float32(hi[c]) - np.float32(lo[c])))   # :289, float range
        out["components"][name] = {"rmse": mean_rmse, "adjusted_loss": adj, "min": lo[c], "max": hi[c]}
    ldir = os.path.join(plt_dir, f"Level_{level}")
    raw_size = float(sum(os.path.getsize(os.path.join(ldir, f)) for f in os.listdir(ldir)))
    raw_size = raw_size / len(hdr.names) * nc                             # src/modes.cpp:316-318
    out["compressed_percent"] = float(sum(xz_sizes)) / raw_size * 100.0   # src/modes.cpp:321-324
    return out


def _dir(d: str) -> str:
    return d if d.endswith("/") else d + "/"      # the reference concatenates dir + name (src/readandwrite.cpp:200)


def compress_run(plt_dirs, min_level: int, max_level: int, components, keep: float, compressed_dir: str,
                 ctx: Context | None = None, threads: int = 0):
    """`-c` (src/modes.cpp:24-112): the five side files, then every (file, level, box, component) in one GPU batch,
    one .xz per unit written by host threads as the chunks come back."""
    ctx = ctx or Context(0)
    compressed_dir = _dir(compressed_dir)
    os.makedirs(compressed_dir, exist_ok=True)
    levels = list(range(min_level, max_level + 1))                         # format_levels
    keys, boxes, dims = [], [], []
    locations, dimensions, counts = [], [], []
    comp_idxs = None
    for t, plt in enumerate(plt_dirs):
        hdr = plotfile.read_header(plt)
        comp_idxs = component_indices(hdr, components)                     # the last file's indices (src/preprocess.cpp:150)
        locations.append([]); dimensions.append([]); counts.append([])
        for li, level in enumerate(levels):
            lev = plotfile.read_level(plt, level)
            locations[t].append([f.lo for f in lev.fabs])
            dimensions[t].append([f.dims for f in lev.fabs])
            counts[t].append(len(lev.fabs))
            for b, fab in enumerate(lev.fabs):
                for c in comp_idxs:
                    keys.append((t, li, c, b))
                    boxes.append(fab.data[c])
                    dims.append(fab.dims)
    runinfo = sidefiles.RunInfo(list(plt_dirs), min_level, max_level, list(components), list(comp_idxs))
    amrexinfo = sidefiles.amrexinfo_from_headers(plt_dirs, len(levels))
    sidefiles.write_all(compressed_dir, runinfo, locations, dimensions, counts, amrexinfo)    # src/modes.cpp:71-89

    def write(i, ser):
        with open(compressed_dir + unit_filename(*keys[i]), "wb") as f:
            f.write(xz_encode(ser))
    plan, rec = _compress_units(ctx, boxes, dims, keep, write, threads, ingest_stats=False)
    plan.close()
    return {"units": len(keys), "npairs": [int(k) for k in rec["npairs"]], "need32": [bool(int(f) & 1) for f in rec["flags"]]}


def decompress_run(compressed_dir: str, out_dir: str, ctx: Context | None = None, threads: int = 0):
    """`-d` (src/modes.cpp:115-204): everything comes from the side files; the unit files are xz-decoded on host
    threads, their pair bytes form one dense stream for a decode plan (float64 out, the widening of
    src/writeplotfile.cpp:103 on the GPU), and each timestep is written as a complete plotfile — Header included."""
    ctx = ctx or Context(0)
    compressed_dir, out_dir = _dir(compressed_dir), _dir(out_dir)
    ri = sidefiles.read_runinfo(compressed_dir)
    nt, nl, nc = len(ri.files), ri.max_level - ri.min_level + 1, len(ri.comp_idxs)
    counts = sidefiles.read_box_counts(compressed_dir, nt, nl)
    locs = sidefiles.read_loc_dim(compressed_dir, "locations.raw", counts)
    dims_ld = sidefiles.read_loc_dim(compressed_dir, "dimensions.raw", counts)
    info = sidefiles.read_amrexinfo(compressed_dir)
    keys = [(t, l, c, b) for t in range(nt) for l in range(nl) for b in range(counts[t][l]) for c in ri.comp_idxs]

    def read(k):
        with open(compressed_dir + unit_filename(*k), "rb") as f:
            raw = xz_decode(f.read())
        h = np.frombuffer(raw, "<i4", 5)
        if len(raw) < 20 or (h < 0).any() or int(h[0]) * int(h[1]) * int(h[2]) != int(h[3]) or h[4] > h[3] or len(raw) - 20 < 8 * int(h[4]):
            raise ValueError(f"corrupt unit file {unit_filename(*k)}")
        return (int(h[0]), int(h[1]), int(h[2])), int(h[4]), raw
    with ThreadPoolExecutor(threads or os.cpu_count() or 1) as pool:
        units = list(pool.map(read, keys))
    for k, (d, _, _) in zip(keys, units):
        if tuple(dims_ld[k[0]][k[1]][k[3]]) != d:
            raise ValueError(f"{unit_filename(*k)}: shape {d} disagrees with dimensions.raw {dims_ld[k[0]][k[1]][k[3]]}")
    npairs = np.array([u[1] for u in units], np.int32)
    stream = np.frombuffer(b"".join(u[2][20:20 + 8 * u[1]] for u in units) or b"\0" * 8, capi.PAIR)
    udims = [u[0] for u in units]
    outs = [np.empty((d[2], d[1], d[0]), np.float64) for d in udims]
    od = capi.box_descs([o.ctypes.data for o in outs], [capi.WC_F64] * len(outs), udims)
    dp = ctx.decode_plan(od, capi.WC_HOST)
    dp.decode(stream.ctypes.data, npairs.ctypes.data, capi.WC_HOST)
    dp.finish()
    dp.close()
    it = iter(outs)
    for t in range(nt):
        name = out_dir + os.path.basename(os.path.normpath(ri.files[t]))       # path.filename(), src/writeplotfile.cpp:132
        level_boxes = []
        for l in range(nl):
            bl = [(tuple(lo), tuple(lo[k] + dm[k] - 1 for k in range(3))) for lo, dm in zip(locs[t][l], dims_ld[t][l])]
            level_boxes.append(bl)
            data = [np.stack([next(it) for _ in range(nc)]) for _ in bl]
            plotfile.write_level(name, l, bl, data, nc)
        plotfile.write_header(name, plotfile.header_from_amrexinfo(ri.components, info, t, level_boxes))
    return ri
