"""Synthetic AMReX-shaped multi-level plotfile data for benchmarks and tests.

"AMR-256-L4" (BASELINE config 3): level 0 = 256^3 domain covered by 64 boxes of 64^3; levels 1-3 =
a centred cube of 256^3 level-l cells (domain (256*2^l)^3, ref_ratio 2) tiled by 512 boxes of 32^3.
16 777 216 cells per level, 8 float64 components -> 4.29 GB, 12 800 units per timestep.

Memory layout mirrors an AMReX FAB file (tests/plt00074/Level_0/Cell_D_00000 in the reference):
per box, the components back to back, each an x-fastest float64 slab — so a unit
(box, component) is one contiguous slab, exactly what wc_box_desc describes.

Values (component c, timestep t, physical position p of the cell centre):
    v = A_c * [ sin(2pi(3 p_x + 0.01 t)) cos(2pi 2 p_y) sin(2pi 5 p_z)
                + tanh((|p - (0.5, 0.5, 0.3 + 0.005 t)| - 0.2) / 0.01) ] + B_c + sigma_c * N(0,1)
The velocity components are sign-symmetric on purpose: they exercise the reference's
negative-max quirk (every coefficient kept, SURVEY.md D3').
The generator is written once against a tiny array-backend shim, so it runs on the GPU (torch,
for the benchmark — never materialised on the host) and on the CPU (numpy, for tests).  The two
backends use different Philox streams; tests that need the same values on both sides copy them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# (A_c, B_c, sigma_c) — density, Temp, pressure, x/y/z velocity, two mass fractions
COMPONENTS = [
    ("density", 0.5, 1.0, 1e-4),
    ("Temp", 800.0, 1500.0, 0.05),
    ("pressure", 2e4, 1e5, 1.0),
    ("x_velocity", 50.0, 0.0, 1e-2),
    ("y_velocity", 50.0, 0.0, 1e-2),
    ("z_velocity", 50.0, 0.0, 1e-2),
    ("Y(H2O)", 1e-3, 1e-3, 1e-7),
    ("Y(CO2)", 1e-3, 1e-3, 1e-7),
]
SEED = 0x5EED0001


@dataclass(frozen=True)
class LevelSpec:
    level: int
    region: int      # cells per side of the covered cube at this level
    box: int         # cells per side of one box
    domain: int      # cells per side of the whole level domain

    @property
    def boxes_per_side(self) -> int:
        return self.region // self.box

    @property
    def n_boxes(self) -> int:
        return self.boxes_per_side ** 3

    @property
    def origin(self) -> int:
        return (self.domain - self.region) // 2


def amr_levels(base: int = 256, n_levels: int = 4, l0_box: int = 64, fine_box: int = 32):
    """Level specs of the AMR-<base>-L<n_levels> layout."""
    out = []
    for l in range(n_levels):
        out.append(LevelSpec(l, base, l0_box if l == 0 else fine_box, base * (1 << l)))
    return out


def _field(xp, lev: LevelSpec, comp: int, t: int, dtype, noise):
    """The (region, region, region) [z][y][x] float64 field of one level / component."""
    A, B, sigma = COMPONENTS[comp % len(COMPONENTS)][1:]
    n = lev.region
    if hasattr(noise, "device"):  # torch: build the coordinates where the noise lives
        g = xp.arange(n, dtype=dtype, device=noise.device) + float(lev.origin)
    else:
        g = xp.arange(n, dtype=dtype) + float(lev.origin)
    p = (g + 0.5) / float(lev.domain)
    px = p.reshape(1, 1, n)
    py = p.reshape(1, n, 1)
    pz = p.reshape(n, 1, 1)
    two_pi = 2.0 * math.pi
    wave = xp.sin(two_pi * (3.0 * px + 0.01 * t)) * xp.cos(two_pi * 2.0 * py) * xp.sin(two_pi * 5.0 * pz)
    r = xp.sqrt((px - 0.5) ** 2 + (py - 0.5) ** 2 + (pz - (0.3 + 0.005 * t)) ** 2)
    front = xp.tanh((r - 0.2) / 0.01)
    return A * (wave + front) + B + sigma * noise


def _to_fab_order(xp, field, lev: LevelSpec):
    """(region^3) [z][y][x] -> (n_boxes, box, box, box): box index z-major, each box x-fastest."""
    b, s = lev.boxes_per_side, lev.box
    v = field.reshape(b, s, b, s, b, s)
    if hasattr(v, "permute"):
        v = v.permute(0, 2, 4, 1, 3, 5)
    else:
        v = v.transpose(0, 2, 4, 1, 3, 5)
    return v.reshape(b * b * b, s, s, s)


def generate_level_numpy(lev: LevelSpec, n_comp: int, t: int = 0, dtype=np.float64) -> np.ndarray:
    """(n_boxes, n_comp, box, box, box) float64, FAB order, on the host."""
    out = np.empty((lev.n_boxes, n_comp, lev.box, lev.box, lev.box), dtype)
    for c in range(n_comp):
        rng = np.random.Generator(np.random.Philox(key=[SEED, (t << 20) | (lev.level << 8) | c]))
        noise = rng.standard_normal((lev.region,) * 3)
        f = _field(np, lev, c, t, np.float64, noise)
        out[:, c] = _to_fab_order(np, f, lev).astype(dtype)
    return out


def generate_level_torch(lev: LevelSpec, n_comp: int, t: int = 0, device="cuda", dtype=None):
    """Same on a torch device (generated there; nothing touches the host)."""
    import torch
    dtype = dtype or torch.float64
    out = torch.empty((lev.n_boxes, n_comp, lev.box, lev.box, lev.box), dtype=dtype, device=device)
    gen = torch.Generator(device=device)
    for c in range(n_comp):
        gen.manual_seed(SEED * 1000003 + ((t << 20) | (lev.level << 8) | c))
        noise = torch.randn((lev.region,) * 3, dtype=torch.float64, device=device, generator=gen)
        f = _field(torch, lev, c, t, torch.float64, noise)
        out[:, c] = _to_fab_order(torch, f, lev).to(dtype)
        del f, noise
    return out


def unit_table(levels, n_comp: int):
    """[(level, box, comp, (nx, ny, nz))] in the reference's iteration order (t, level, box) with the
    components of a box adjacent (src/iterator.h:24-33, src/compressor.cpp:203)."""
    units = []
    for lev in levels:
        for b in range(lev.n_boxes):
            for c in range(n_comp):
                units.append((lev.level, b, c, (lev.box, lev.box, lev.box)))
    return units


def shard_units(n_units_or_sizes, world: int, rank: int):
    """Static size-balanced contiguous partition of a unit list over `world` ranks:
    returns the [start, stop) slice of rank `rank`.  Units of one box stay together when the caller
    passes per-box sizes."""
    sizes = np.asarray(n_units_or_sizes, dtype=np.int64)
    if sizes.ndim == 0:
        sizes = np.ones(int(sizes), np.int64)
    cum = np.concatenate([[0], np.cumsum(sizes)])
    total = cum[-1]
    bounds = [int(np.searchsorted(cum, total * r / world, side="left")) for r in range(world + 1)]
    bounds[0], bounds[-1] = 0, len(sizes)
    return bounds[rank], bounds[rank + 1]
