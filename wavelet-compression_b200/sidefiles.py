"""The five .raw side files of a compression run, byte-compatible with the reference's writers / readers
(src/readandwrite.cpp:226-395; written by `-c` at src/modes.cpp:71-89, read by `-d` at :117-181):

    runinfo.raw     vector<string> files | int min_level | int max_level | vector<string> components | vector<int> comp_idxs
    locations.raw   3 floats (lo.x, lo.y, lo.z) per box, in (timestep, level, box) order       (ints stored AS float32)
    dimensions.raw  3 floats (nx, ny, nz) per box, same order
    boxcounts.raw   one float per (timestep, level): the number of boxes
    amrexinfo.raw   vector<vector<double>> geomcellinfo | vector<int> ref_ratios | vector<long double> true_times |
                    vector<vector<int>> level_steps | int xDim | int yDim | int zDim

Encoding (native x86-64, as `stream.write(reinterpret_cast<const char*>(&v), sizeof v)` produces): size_t = 8 bytes LE,
int = 4 bytes LE, float / double IEEE LE, long double = 80-bit x87 extended in a 16-byte slot.  The 6 padding bytes of
a long double slot are indeterminate in the reference (they are whatever the stack held); this writer zeroes them and
readers / comparisons must ignore them (`mask_long_double_padding`).

These files are metadata, not the hot path: nothing here touches the GPU.  `amrexinfo_from_headers` reproduces what
src/preprocess.cpp:166-259 extracts from the plotfile `Header`s, quirks included.
"""
from __future__ import annotations

import os
import re
import struct
from dataclasses import dataclass, field

import numpy as np

NAMES = ("runinfo.raw", "locations.raw", "dimensions.raw", "boxcounts.raw", "amrexinfo.raw")


@dataclass
class RunInfo:                      # src/box-structs.h:22-28
    files: list
    min_level: int
    max_level: int
    components: list
    comp_idxs: list


@dataclass
class AMReXInfo:                    # src/box-structs.h:42-50
    geomcellinfo: list              # per timestep [lo.x lo.y lo.z hi.x hi.y hi.z]
    ref_ratios: list                # 3 ints
    true_times: list                # per timestep: decimal TEXT (parsed to long double on write) or np.longdouble
    level_steps: list               # per timestep, one int per selected level
    xDim: int = 0
    yDim: int = 0
    zDim: int = 0
    _times_ld: list = field(default_factory=list, repr=False)


# ---- primitives ---------------------------------------------------------------------------------
def _size(n):
    return struct.pack("<Q", n)


def _int(v):
    return struct.pack("<i", int(v))


def _string(s):
    b = s.encode()
    return _size(len(b)) + b


def _vec_string(v):
    return _size(len(v)) + b"".join(_string(s) for s in v)


def _vec_int(v):
    return _size(len(v)) + b"".join(_int(x) for x in v)


def _long_double(x) -> bytes:
    assert np.dtype(np.longdouble).itemsize == 16, "x86-64 long double expected"
    ld = np.longdouble(x) if not isinstance(x, np.longdouble) else x      # text -> strtold precision, as operator>>
    return np.array([ld], np.longdouble).tobytes()[:10] + b"\0" * 6


class _Reader:
    def __init__(self, buf: bytes):
        self.b, self.o = buf, 0

    def take(self, n):
        if self.o + n > len(self.b):
            raise ValueError("side file truncated")
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def size(self):
        return struct.unpack("<Q", self.take(8))[0]

    def int(self):
        return struct.unpack("<i", self.take(4))[0]

    def string(self):
        return self.take(self.size()).decode()

    def vec_string(self):
        return [self.string() for _ in range(self.size())]

    def vec_int(self):
        return [self.int() for _ in range(self.size())]

    def long_double(self):
        raw = self.take(16)
        return np.frombuffer(raw[:10] + b"\0" * 6, np.longdouble)[0]


def mask_long_double_padding(amrexinfo_bytes: bytes) -> bytes:
    """amrexinfo.raw with the 6 padding bytes of every long double slot zeroed (they are indeterminate)."""
    r = _Reader(amrexinfo_bytes)
    for _ in range(r.size()):
        r.take(8 * r.size())
    r.take(4 * r.size())
    n = r.size()
    out = bytearray(amrexinfo_bytes)
    for i in range(n):
        out[r.o + 16 * i + 10:r.o + 16 * i + 16] = b"\0" * 6
    return bytes(out)


# ---- writers (src/readandwrite.cpp:226-376) ---------------------------------------------------------
def _path(d, name):
    return d + name            # the reference concatenates: the directory must end with '/' (src/readandwrite.cpp:200)


def write_runinfo(d: str, info: RunInfo):
    with open(_path(d, "runinfo.raw"), "wb") as f:
        f.write(_vec_string(info.files) + _int(info.min_level) + _int(info.max_level) + _vec_string(info.components) +
                _vec_int(info.comp_idxs))


def write_loc_dim(d: str, name: str, data):
    """data[t][level][box] = (a, b, c) ints, written as float32 (write_float(file, value) with an int value)."""
    flat = [float(v) for t in data for lev in t for box in lev for v in box]
    with open(_path(d, name), "wb") as f:
        f.write(np.asarray(flat, "<f4").tobytes())


def write_box_counts(d: str, counts):
    with open(_path(d, "boxcounts.raw"), "wb") as f:
        f.write(np.asarray([float(c) for t in counts for c in t], "<f4").tobytes())


def write_amrexinfo(d: str, info: AMReXInfo):
    out = _size(len(info.geomcellinfo))
    for g in info.geomcellinfo:
        out += _size(len(g)) + np.asarray(g, "<f8").tobytes()
    out += _vec_int(info.ref_ratios)
    out += _size(len(info.true_times)) + b"".join(_long_double(t) for t in info.true_times)
    out += _size(len(info.level_steps))
    for ls in info.level_steps:
        out += _vec_int(ls)
    out += _int(info.xDim) + _int(info.yDim) + _int(info.zDim)
    with open(_path(d, "amrexinfo.raw"), "wb") as f:
        f.write(out)


# ---- readers (src/readandwrite.cpp:245-395) ---------------------------------------------------------
def read_runinfo(d: str) -> RunInfo:
    r = _Reader(open(_path(d, "runinfo.raw"), "rb").read())
    files = r.vec_string()
    lo, hi = r.int(), r.int()
    return RunInfo(files, lo, hi, r.vec_string(), r.vec_int())


def read_box_counts(d: str, num_times: int, num_levels: int):
    v = np.frombuffer(open(_path(d, "boxcounts.raw"), "rb").read(), "<f4")
    if v.size < num_times * num_levels:
        raise ValueError("boxcounts.raw truncated")
    return [[int(v[t * num_levels + l]) for l in range(num_levels)] for t in range(num_times)]


def read_loc_dim(d: str, name: str, counts):
    v = np.frombuffer(open(_path(d, name), "rb").read(), "<f4")
    need = 3 * sum(sum(t) for t in counts)
    if v.size < need:
        raise ValueError(name + " truncated")
    out, k = [], 0
    for t in counts:
        row = []
        for c in t:
            row.append([tuple(int(x) for x in v[k + 3 * b:k + 3 * b + 3]) for b in range(c)])
            k += 3 * c
        out.append(row)
    return out


def read_amrexinfo(d: str) -> AMReXInfo:
    r = _Reader(open(_path(d, "amrexinfo.raw"), "rb").read())
    geom = [list(np.frombuffer(r.take(8 * r.size()), "<f8")) for _ in range(r.size())]
    ref = r.vec_int()
    times = [r.long_double() for _ in range(r.size())]
    steps = [r.vec_int() for _ in range(r.size())]
    return AMReXInfo([[float(x) for x in g] for g in geom], ref, times, steps, r.int(), r.int(), r.int())


# ---- what preprocess_data extracts from the Headers (src/preprocess.cpp:166-259) -----------------------
def amrexinfo_from_headers(plt_dirs, num_levels: int) -> AMReXInfo:
    """One pass over each plotfile's text Header, mirroring the reference's stream extraction:
      * true_time: the time line parsed as long double (operator>>)                     :183-186
      * geomcell: three doubles of the prob_lo line, three of the prob_hi line           :190-207
      * ref_ratios: from the FIRST file only, `dim` ints off the ref-ratio line; a missing token reads as 0
        (C++11 failed extraction), so a 2-level file gives [2, 0, 0]                     :210-221
      * xDim,yDim,zDim: the numbers after the third '(' of the domain line, + 1 (last file wins)   :224-245
      * level_steps: the first `num_levels` ints of the level-steps line                   :249-257
    """
    info = AMReXInfo([], [], [], [])
    for i, plt in enumerate(plt_dirs):
        lines = open(os.path.join(plt, "Header")).read().split("\n")
        ncomp = int(lines[1])
        k = 2 + ncomp
        dim = int(lines[k].split()[0])
        info.true_times.append(lines[k + 1].split()[0])
        lo = [float(v) for v in lines[k + 3].split()[:3]]
        hi = [float(v) for v in lines[k + 4].split()[:3]]
        info.geomcellinfo.append(lo + hi)
        if i == 0:
            toks = lines[k + 5].split()
            info.ref_ratios = [int(toks[j]) if j < len(toks) else 0 for j in range(dim)]
        dom = lines[k + 6]
        p = -1
        for _ in range(3):
            p = dom.find("(", p + 1)
        e = dom.find(")", p)
        d3 = [int(re.match(r"\s*-?\d+", v).group(0)) for v in dom[p + 1:e + 1].rstrip(")").split(",")]
        info.xDim, info.yDim, info.zDim = d3[0] + 1, d3[1] + 1, d3[2] + 1
        toks = lines[k + 7].split()
        info.level_steps.append([int(toks[j]) if j < len(toks) else 0 for j in range(num_levels)])
    return info


def write_all(d: str, runinfo: RunInfo, locations, dimensions, counts, amrexinfo: AMReXInfo):
    """The five writes of src/modes.cpp:71-89, in that order."""
    write_runinfo(d, runinfo)
    write_loc_dim(d, "locations.raw", locations)
    write_loc_dim(d, "dimensions.raw", dimensions)
    write_box_counts(d, counts)
    write_amrexinfo(d, amrexinfo)
