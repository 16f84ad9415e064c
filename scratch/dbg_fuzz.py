import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as g
import torch
from conftest import smooth_box, same_bits
from test_gpu_fuzz import _shapes
from oracle.pyoracle import Oracle
pkg = g.package(); orc = Oracle(); wc = pkg
ctx = pkg.Context(0)
seed = int(sys.argv[1])
rng = np.random.default_rng(seed)
shapes = _shapes(rng, 70)
shapes += [(8, 8, 8)] * 3 + [(16, 16, 16)] * 3 + [(32, 32, 32)] * 2 + [(64, 64, 64)] + [(8, 8, 4), (16, 16, 12), (32, 32, 28), (64, 64, 60)]
boxes, dts = [], []
for i, d in enumerate(shapes):
    dt = np.float32 if rng.random() < 0.3 else np.float64
    b = smooth_box(d, rng, dtype=dt, sym=bool(i % 2), noise=10.0 ** -int(rng.integers(0, 6)))
    if rng.random() < 0.08: b = -np.abs(b) - 0.5
    if rng.random() < 0.05: b = np.zeros_like(b)
    boxes.append(b); dts.append(dt)
keep = float(np.float32(rng.choice([0.9, 0.99, 0.999, 0.9999])))
dev = [torch.from_numpy(np.ascontiguousarray(b)).cuda() for b in boxes]
outs = [torch.full((int(np.prod(d)),), 3.0, dtype=torch.float32, device="cuda") for d in shapes]
torch.cuda.synchronize()
code = lambda dt: wc.WC_F64 if dt == np.float64 else wc.WC_F32
descs = wc.capi.box_descs([t.data_ptr() for t in dev], [code(dt) for dt in dts], shapes)
odescs = wc.capi.box_descs([t.data_ptr() for t in outs], [wc.WC_F32] * len(outs), shapes)
plan = ctx.plan(descs, wc.WC_DEVICE)
for k2 in (keep, float(np.float32(0.95))):
    plan.compress(k2); plan.decompress(odescs, wc.WC_DEVICE); ctx.sync()
    got = plan.fetch_host()
    for i, (b, d) in enumerate(zip(boxes, shapes)):
        runs, vals, _ = orc.compress_unit(b, d, k2)
        okp = same_bits(got[i].runs, runs) and same_bits(got[i].vals, vals)
        ob = orc.decompress_unit(runs, vals, d).reshape(-1)
        r = outs[i].cpu().numpy()
        bad = np.nonzero((r.view(np.uint32) != ob.view(np.uint32)) & ~(np.isnan(r) & np.isnan(ob)))[0]
        if not okp or bad.size:
            n = int(np.prod(d))
            print('keep', k2, 'unit', i, d, dts[i].__name__, 'N', n, 'pairs ok', okp, 'K', runs.size, 'mismatch', bad.size, bad[:6], 'in ptr%16', dev[i].data_ptr() % 16, 'out ptr%16', outs[i].data_ptr() % 16)
print('done', len(shapes))
