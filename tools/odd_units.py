"""Per-kernel breakdown on boxes that stay on the generic kernels (40x40x42: nz % 4 != 0)."""
import sys
sys.path.insert(0, '.')
import __graft_entry__ as g
import torch
pkg = g.package()
stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream=stream.cuda_stream)
KEEP = 0.9990000128746033
dims, n_units = (40, 40, 42), 16384
n = dims[0] * dims[1] * dims[2]
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
