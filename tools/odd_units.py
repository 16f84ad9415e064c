"""Per-kernel breakdown on boxes the y-slab classes refuse (40x40x42: nz % 4 != 0; odd dimensions): the x-slab classes of
wc_xslab.cu (path 0) against the generic multi-kernel path (path 1).
usage: python tools/odd_units.py [X Y Z [n_units [path]]]"""
import sys
sys.path.insert(0, '.')
import __graft_entry__ as g
import torch
pkg = g.package()
stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream=stream.cuda_stream)
KEEP = 0.9990000128746033
a = [int(v) for v in sys.argv[1:]]
dims = tuple(a[:3]) if len(a) >= 3 else (40, 40, 42)
n = dims[0] * dims[1] * dims[2]
n_units = a[3] if len(a) >= 4 and a[3] > 0 else max(1, (1 << 30) // (8 * n))      # ~1 GB of float64 input
path = a[4] if len(a) >= 5 else 0
ctx.set_path(path)
print("dims", dims, "units", n_units, "path", path, "(0 = auto: x-slab classes, 1 = generic kernels)")
gen = torch.Generator(device='cuda'); gen.manual_seed(1)
x = torch.linspace(0, 50, n_units * n, device='cuda', dtype=torch.float64).sin_() * 100 + \
    torch.randn(n_units * n, device='cuda', dtype=torch.float64, generator=gen) * 0.05
descs = pkg.capi.box_descs([x.data_ptr() + 8 * n * i for i in range(n_units)], [pkg.WC_F64] * n_units, [dims] * n_units)
rec = torch.empty(n_units * n, dtype=torch.float32, device='cuda')
odescs = pkg.capi.box_descs([rec.data_ptr() + 4 * n * i for i in range(n_units)], [pkg.WC_F32] * n_units, [dims] * n_units)
torch.cuda.synchronize()
plan = ctx.plan(descs, pkg.WC_DEVICE)
with torch.cuda.stream(stream):
    for _ in range(2): plan.compress(KEEP); plan.decompress(odescs, pkg.WC_DEVICE)
    torch.cuda.synchronize()
    ctx.set_profile(True); ctx.reset_counters()
    for _ in range(3): plan.compress(KEEP)
    ctx.sync()
    print("compress kernels (ms per step):")
    tot = 0
    for k, (c, ms) in ctx.kernel_stats().items(): print("  ", k, c // 3, round(ms / 3, 3)); tot += ms / 3
    K = plan.total_pairs()
    print("  total", round(tot, 3), "ms; alg", round((8 * n * n_units + 8 * K) / tot / 1e6), "GB/s; kept", K / (n * n_units))
    ctx.reset_counters()
    for _ in range(3): plan.decompress(odescs, pkg.WC_DEVICE)
    ctx.sync()
    print("decompress kernels (ms per step):")
    tot = 0
    for k, (c, ms) in ctx.kernel_stats().items(): print("  ", k, c // 3, round(ms / 3, 3)); tot += ms / 3
    print("  total", round(tot, 3), "ms; alg", round((8 * K + 4 * n * n_units) / tot / 1e6), "GB/s")
