"""AMReX plotfile I/O without AMReX: just enough of the on-disk format
(`Header`, `Level_k/Cell_H`, `Level_k/Cell_D_*`) to feed the GPU path with raw FAB payloads and to write
decompressed boxes back.  Format as found in the reference's fixtures (tests/plt00074) and read there by
amrex::VisMF::Read (src/preprocess.cpp:36) / written by WriteMultiLevelPlotfile (src/writeplotfile.cpp:220):

  Cell_H : "1\n1\n<ncomp>\n<nghost>\n(<nboxes> 0\n((lo) (hi) (0,0,0))\n...)\n<nfabs>\n
            FabOnDisk: <file> <offset>\n...\n\n<nboxes>,<ncomp>\n<min table>\n\n<nboxes>,<ncomp>\n<max table>\n\n"
  Cell_D : per FAB an ASCII line "FAB ((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))((lo) (hi) (0,0,0)) ncomp\n"
           followed by ncomp*N native little-endian float64, component-major, x fastest.

A unit (box, component) is therefore one contiguous float64 slab inside the file image: `level_descs()` returns
wc_box_desc records pointing straight into the (optionally pinned) buffer the file was read into — the narrowing
to float32 (src/preprocess.cpp:78) happens on the GPU.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field

import numpy as np

FAB_REAL_DESC = "((8, (64 11 52 0 1 12 0 1023)),(8, (8 7 6 5 4 3 2 1)))"
_BOX_RE = re.compile(r"\(\((-?\d+),(-?\d+),(-?\d+)\) \((-?\d+),(-?\d+),(-?\d+)\) \((-?\d+),(-?\d+),(-?\d+)\)\)")


@dataclass
class Fab:
    lo: tuple
    hi: tuple
    ncomp: int
    data: np.ndarray          # float64 [ncomp][nz][ny][nx] (a view into the file image)

    @property
    def dims(self):
        return tuple(h - l + 1 for l, h in zip(self.lo, self.hi))


@dataclass
class Level:
    ncomp: int
    fabs: list = field(default_factory=list)


@dataclass
class Header:
    version: str
    names: list
    dim: int
    time: float
    finest_level: int
    prob_lo: list
    prob_hi: list
    ref_ratio: list
    domains: list             # per level ((lo), (hi))
    level_steps: list
    cell_size: list           # per level [dx, dy, dz]
    coord_sys: int
    bwidth: int
    levels: list              # per level dict(level, ngrids, time, step, boxes_phys [[(lo,hi) x3] per grid], path)


def _fmt(x: float) -> str:
    """AMReX writes reals with 17 significant digits."""
    return "%.17g" % x


def read_header(plt_dir: str) -> Header:
    """Parses the text `Header` (what src/preprocess.cpp:128-258 extracts with ifstream >>)."""
    tok = open(os.path.join(plt_dir, "Header")).read().split("\n")
    it = iter(tok)
    version = next(it).strip()
    ncomp = int(next(it))
    names = [next(it).strip() for _ in range(ncomp)]
    dim = int(next(it))
    time = float(next(it))
    finest = int(next(it))
    prob_lo = [float(v) for v in next(it).split()]
    prob_hi = [float(v) for v in next(it).split()]
    ref_ratio = [int(v) for v in next(it).split()]
    dom_line = next(it)
    domains = [((int(m[0]), int(m[1]), int(m[2])), (int(m[3]), int(m[4]), int(m[5]))) for m in _BOX_RE.findall(dom_line)]
    level_steps = [int(v) for v in next(it).split()]
    cell_size = [[float(v) for v in next(it).split()] for _ in range(finest + 1)]
    coord_sys = int(next(it))
    bwidth = int(next(it))
    levels = []
    for _ in range(finest + 1):
        lev, ngrids, ltime = next(it).split()
        step = int(next(it))
        boxes = []
        for _g in range(int(ngrids)):
            boxes.append([tuple(float(v) for v in next(it).split()) for _d in range(dim)])
        path = next(it).strip()
        levels.append(dict(level=int(lev), ngrids=int(ngrids), time=float(ltime), step=step, boxes_phys=boxes, path=path))
    return Header(version, names, dim, time, finest, prob_lo, prob_hi, ref_ratio, domains, level_steps, cell_size,
                  coord_sys, bwidth, levels)


def write_header(plt_dir: str, h: Header) -> None:
    out = [h.version, str(len(h.names))] + list(h.names) + [str(h.dim), _fmt(h.time), str(h.finest_level)]
    out.append(" ".join(_fmt(v) for v in h.prob_lo) + " ")
    out.append(" ".join(_fmt(v) for v in h.prob_hi) + " ")
    out.append(" ".join(str(v) for v in h.ref_ratio) + " ")
    out.append(" ".join("((%d,%d,%d) (%d,%d,%d) (0,0,0))" % (*lo, *hi) for lo, hi in h.domains) + " ")
    out.append(" ".join(str(v) for v in h.level_steps) + " ")
    for cs in h.cell_size:
        out.append(" ".join(_fmt(v) for v in cs) + " ")
    out += [str(h.coord_sys), str(h.bwidth)]
    for lv in h.levels:
        out.append("%d %d %s" % (lv["level"], lv["ngrids"], _fmt(lv["time"])))
        out.append(str(lv["step"]))
        for g in lv["boxes_phys"]:
            for d in g:
                out.append(" ".join(_fmt(v) for v in d))
        out.append(lv["path"])
    os.makedirs(plt_dir, exist_ok=True)
    with open(os.path.join(plt_dir, "Header"), "w") as f:
        f.write("\n".join(out) + "\n")


def header_from_amrexinfo(names, info, t: int, level_boxes, level_dirs=None) -> Header:
    """The Header amrex::WriteMultiLevelPlotfile writes for timestep `t` of a decompressed run, from the quantities the
    reference carries in amrexinfo.raw (src/writeplotfile.cpp:138-227 builds the Geometry objects, AMReX's
    WriteGenericPlotfileHeader prints them).  Same arithmetic, in the same order, in float64:
      domain of level l   = [0, dim * ref^l - 1]                                  (writeplotfile.cpp:166-172)
      cell size           = (prob_hi - prob_lo) / n_cells                         (Geometry::define)
      box extents         = prob_lo + dx * lo  ..  prob_lo + dx * (hi + 1)        (RealBox(box, dx, prob_lo))
    level_boxes[l] = [(lo, hi)] in index space.  `info` is a sidefiles.AMReXInfo."""
    nlev = len(level_boxes)
    g = info.geomcellinfo[t]
    prob_lo, prob_hi = [float(v) for v in g[:3]], [float(v) for v in g[3:6]]
    time = float(np.longdouble(info.true_times[t]))       # amrex::Real time = long double -> double
    dims0 = [info.xDim, info.yDim, info.zDim]
    # DEVIATION (documented): the reference fills ref_ratios by reading `dim` ints off a Header line that holds one
    # ratio per coarser LEVEL (src/preprocess.cpp:210-221), so a 2-level plotfile yields [2, 0, 0] and its own `-d` would
    # build an empty domain for level 1 (yDim * pow(0, 1)).  The side file keeps the reference's bytes; here a zero
    # entry falls back to the first ratio (AMReX plotfiles refine isotropically), which is what its writer test feeds
    # in directly (src/writeplotfile.cpp:371: {2, 2, 2}).
    ratios = [int(r) if int(r) > 0 else int(info.ref_ratios[0]) for r in info.ref_ratios]
    domains, cell_size, levels = [], [], []
    for l in range(nlev):
        n = [int(dims0[k] * pow(float(ratios[k]), l)) for k in range(3)]               # int(xDim * pow(ref, l))
        domains.append(((0, 0, 0), (n[0] - 1, n[1] - 1, n[2] - 1)))
        dx = [(prob_hi[k] - prob_lo[k]) / float(n[k]) for k in range(3)]
        cell_size.append(dx)
        phys = [[(prob_lo[k] + dx[k] * float(lo[k]), prob_lo[k] + dx[k] * float(hi[k] + 1)) for k in range(3)]
                for lo, hi in level_boxes[l]]
        levels.append(dict(level=l, ngrids=len(level_boxes[l]), time=time, step=int(info.level_steps[t][l]),
                           boxes_phys=phys, path=(level_dirs[l] if level_dirs else f"Level_{l}/Cell")))
    # one ref_ratio entry per coarser level; AMReX prints ref_ratio[i][0]
    return Header("HyperCLaw-V1.1", list(names), 3, time, nlev - 1, prob_lo, prob_hi,
                  [ratios[0]] * (nlev - 1), domains, [int(v) for v in info.level_steps[t][:nlev]], cell_size,
                  0, 0, levels)


def read_level(plt_dir: str, level: int, alloc=None) -> Level:
    """Reads Level_<level>: Cell_H for the box list and FAB offsets, then the Cell_D file images.
    `alloc(nbytes) -> writable uint8 ndarray` lets the caller supply pinned memory (wc_host_alloc)."""
    ldir = os.path.join(plt_dir, f"Level_{level}")
    lines = open(os.path.join(ldir, "Cell_H")).read().split("\n")
    ncomp = int(lines[2])
    nghost = int(lines[3].split()[0].strip("(),") or 0) if lines[3].strip() else 0
    if nghost != 0:
        # the on-disk FAB of a grown MultiFab covers the grown box; the reference reads the valid box out of it
        # (mfi.validbox(), src/preprocess.cpp:43).  Plotfiles are written without ghost cells; refuse rather than
        # hand mis-strided slabs to the GPU.
        raise ValueError(f"{ldir}/Cell_H: nghost = {nghost}; only plotfiles without ghost cells are supported")
    nbox = int(re.match(r"\((\d+)", lines[4]).group(1))
    boxes = []
    for ln in lines[5:5 + nbox]:
        m = _BOX_RE.search(ln)
        boxes.append((tuple(int(m.group(i)) for i in (1, 2, 3)), tuple(int(m.group(i)) for i in (4, 5, 6))))
    fods = [ln.split() for ln in lines if ln.startswith("FabOnDisk:")]
    images = {}
    lev = Level(ncomp)
    for (lo, hi), (_, fname, off) in zip(boxes, fods):
        if fname not in images:
            path = os.path.join(ldir, fname)
            size = os.path.getsize(path)
            buf = alloc(size) if alloc else np.empty(size, np.uint8)
            with open(path, "rb") as f:
                f.readinto(memoryview(buf)[:size])
            images[fname] = buf
        img = images[fname]
        off = int(off)
        nl = off + bytes(img[off:off + 1024]).index(b"\n") + 1    # end of the ASCII "FAB ..." line
        fab_line = bytes(img[off:nl]).decode("ascii", "replace")
        fm = _BOX_RE.search(fab_line)
        if not fm or (tuple(int(fm.group(i)) for i in (1, 2, 3)), tuple(int(fm.group(i)) for i in (4, 5, 6))) != (lo, hi):
            raise ValueError(f"{ldir}/{fname}@{off}: FAB header box {fab_line.strip()!r} does not match the Cell_H box {lo}-{hi}")
        if int(fab_line.split()[-1]) != ncomp:
            raise ValueError(f"{ldir}/{fname}@{off}: FAB has {fab_line.split()[-1]} components, Cell_H says {ncomp}")
        dims = tuple(h - l + 1 for l, h in zip(lo, hi))
        n = dims[0] * dims[1] * dims[2]
        raw = img[nl:nl + 8 * n * ncomp]
        if raw.ctypes.data % 8:
            # the payload starts right after a text line, so it is rarely 8-byte aligned in the file image;
            # realign inside the same buffer when there is slack, else copy (the GPU path wants 16 B)
            raw = np.frombuffer(bytes(raw), np.uint8)
        data = raw.view("<f8").reshape(ncomp, dims[2], dims[1], dims[0])
        lev.fabs.append(Fab(lo, hi, ncomp, data))
    return lev


def level_units(lev: Level, comp_idxs):
    """[(slab float64 [nz][ny][nx], (nx,ny,nz))] in the reference's (box, component) order
    (src/iterator.h:24-33, src/compressor.cpp:203)."""
    units = []
    for fab in lev.fabs:
        for c in comp_idxs:
            units.append((fab.data[c], fab.dims))
    return units


def write_level(plt_dir: str, level: int, boxes, data, ncomp: int, nghost: int = 0) -> None:
    """Writes Level_<level>/Cell_D_00000 + Cell_H.  boxes: [(lo, hi)], data: per box float64 [ncomp][nz][ny][nx]
    (float32 input is widened, as src/writeplotfile.cpp:103 does)."""
    ldir = os.path.join(plt_dir, f"Level_{level}")
    os.makedirs(ldir, exist_ok=True)
    offsets, mins, maxs = [], [], []
    with open(os.path.join(ldir, "Cell_D_00000"), "wb") as f:
        for (lo, hi), d in zip(boxes, data):
            d = np.ascontiguousarray(d, dtype="<f8")
            offsets.append(f.tell())
            f.write(("FAB %s((%d,%d,%d) (%d,%d,%d) (0,0,0)) %d\n" % (FAB_REAL_DESC, *lo, *hi, ncomp)).encode())
            f.write(d.tobytes())
            mins.append([float(d[c].min()) for c in range(ncomp)])
            maxs.append([float(d[c].max()) for c in range(ncomp)])
    out = ["1", "1", str(ncomp), str(nghost), "(%d 0" % len(boxes)]
    out += ["((%d,%d,%d) (%d,%d,%d) (0,0,0))" % (*lo, *hi) for lo, hi in boxes]
    out += [")", str(len(boxes))]
    out += ["FabOnDisk: Cell_D_00000 %d" % o for o in offsets]
    out.append("")
    for table in (mins, maxs):
        out.append("%d,%d" % (len(boxes), ncomp))
        for row in table:
            out.append("".join("%.16e," % v for v in row))
        out.append("")
    with open(os.path.join(ldir, "Cell_H"), "w") as f:
        f.write("\n".join(out) + "\n")
