"""Probe: one / many 64^3 units through the cluster compress kernel of the library named by WCGPU_LIB, against the oracle."""
import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
from conftest import same_bits, smooth_box
from oracle.pyoracle import Oracle
pkg = g.package(); orc = Oracle()
rng = np.random.default_rng(5)
keep = float(np.float32(0.999))
ctx = pkg.Context(0); ctx.set_path(2)
for n, d in ((1, (64, 64, 64)), (40, (64, 64, 64)), (3, (40, 40, 40)), (3, (36, 36, 36)), (3, (48, 48, 48))):
    boxes = [smooth_box(d, rng, dtype=np.float64) for _ in range(n)]
    t = time.time()
    packed = ctx.compress_batch(boxes, keep, dims=[d] * n)
    ok = True
    for b, p in zip(boxes[:4], packed[:4]):
        runs, vals, _ = orc.compress_unit(b, d, keep)
        ok &= p.npairs == runs.size and same_bits(p.runs, runs) and same_bits(p.vals, vals)
    print("units", n, d, "ok" if ok else "MISMATCH", f"{time.time()-t:.2f}s", flush=True)
