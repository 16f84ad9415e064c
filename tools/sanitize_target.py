"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): a few units of every
kernel class (fused cubes, clusters of 2 / 4 / 8, small units, big boxes by y-slabs, odd dimensions on the x-slab
kernels, one box on the generic kernels) through compress -> plan round trip -> stream decode (staged and direct kernels, both segment index kernels)
-> RMSE -> unit stats, checked against the oracle.  Kept tiny: kernels run 10-100x slower under the tool.

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py
"""
import sys

import numpy as np

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import __graft_entry__ as g
from conftest import same_bits, smooth_box
from oracle.pyoracle import Oracle

pkg = g.package()
capi = pkg.capi
orc = Oracle()
rng = np.random.default_rng(3)
dims = ([(32, 32, 32)] * 3 + [(64, 64, 64)] * 2 + [(16, 16, 16)] * 3 + [(8, 8, 8)] * 3 + [(40, 40, 40)] + [(36, 36, 36)] +
        [(16, 32, 64)] * 2 + [(8, 4, 4)] + [(5, 7, 3)] + [(8, 4, 2)] + [(44, 44, 44)] + [(64, 72, 64)] + [(34, 18, 10)] +
        [(31, 17, 9), (33, 33, 33), (40, 40, 42), (63, 47, 41), (65, 61, 57)] +     # x-slab classes: one CTA, clusters of 2 / 4 / 8
        [(96, 96, 50)])                                                             # generic kernels (no fused class holds it)
boxes = [smooth_box(d, rng, dtype=np.float64 if i % 2 else np.float32, sym=(i % 3 == 0)) for i, d in enumerate(dims)]
keep = float(np.float32(0.999))
ctx = pkg.Context(0)
ctx.set_option(capi.WC_OPT_INGEST_STATS, 1)
ok = True
# 1. compress (fused + generic classes, cluster kernels, min/max instantiations) and the plan round trip
plan = ctx.plan_host(boxes, dims)
plan.compress(keep)
got = plan.fetch_host()
lo, hi, n32 = plan.unit_stats()
want = [orc.compress_unit(b, d, keep) for b, d in zip(boxes, dims)]
for i, (p, (runs, vals, _)) in enumerate(zip(got, want)):
    ok &= same_bits(p.runs, runs) and same_bits(p.vals, vals)
    ok &= bool(lo[i] == np.nanmin(boxes[i].astype(np.float32)) and hi[i] == np.nanmax(boxes[i].astype(np.float32)))
recon = [np.zeros((d[2], d[1], d[0]), np.float32) for d in dims]
od = capi.box_descs([r.ctypes.data for r in recon], [capi.WC_F32] * len(dims), dims)
for pipe in (1, 2, 0):
    ctx.set_option(capi.WC_OPT_DECODE_PIPE, pipe)
    for r in recon:
        r[:] = 7
    plan.decompress(od, capi.WC_HOST)
    ctx.sync()
    for i, (r, (runs, vals, _)) in enumerate(zip(recon, want)):
        ok &= same_bits(r, orc.decompress_unit(runs, vals, dims[i]))
plan.close()
# 2. global threshold mode (keys-only + given-threshold kernels)
packed_g = ctx.compress_batch(boxes, keep, thresh_mode=capi.WC_THRESH_GLOBAL, dims=dims)
flats = [orc.haar_forward(orc.narrow(b) if b.dtype == np.float64 else b, d) for b, d in zip(boxes, dims)]
tg = orc.select_threshold_global(flats, keep)
for p, f in zip(packed_g, flats):
    rg, vg = orc.threshold_pack(f, tg)
    ok &= same_bits(p.runs, rg) and same_bits(p.vals, vg)
# 3. the stream (-d) path: dense pairs + counts through decode plans
k = np.array([w[0].size for w in want], np.int32)
pr = np.empty(max(int(k.sum()), 1), capi.PAIR)
o = 0
for runs, vals, _ in want:
    pr["run"][o:o + runs.size], pr["val"][o:o + runs.size] = runs, vals
    o += runs.size
for seg, pipe in ((0, 1), (0, 2), (1, 0), (0, 0)):
    ctx.set_option(capi.WC_OPT_SEG_INDEX, seg)
    ctx.set_option(capi.WC_OPT_DECODE_PIPE, pipe)
    for r in recon:
        r[:] = 7
    dp = ctx.decode_plan(od, capi.WC_HOST)
    dp.decode(pr.ctypes.data, k.ctypes.data, capi.WC_HOST)
    dp.finish()
    dp.close()
    for i, (r, (runs, vals, _)) in enumerate(zip(recon, want)):
        ok &= same_bits(r, orc.decompress_unit(runs, vals, dims[i]))
# 4. RMSE
b32 = [b.astype(np.float32) for b in boxes]
rm = ctx.rmse_batch(b32, recon)
for i, d in enumerate(dims):
    oe = orc.rmse(b32[i], recon[i], d)
    ok &= abs(rm[i] - oe) <= 1e-12 * max(abs(oe), 1e-300)
ctx.close()
print("sanitize_target", "ok" if ok else "MISMATCH", "units", len(dims))
sys.exit(0 if ok else 1)
