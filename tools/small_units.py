"""Throughput of the fused path on many small units (16^3 / 8^3 / 16x32x64 boxes): not a BASELINE config,
just to know where the runtime-geometry kernels stand."""
import sys
sys.path.insert(0, '.')
import __graft_entry__ as g
import torch
pkg = g.package()
stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream=stream.cuda_stream)
KEEP = 0.9990000128746033
for dims, n_units in (((16, 16, 16), 262144), ((8, 8, 8), 1048576), ((16, 32, 64), 32768), ((32, 32, 32), 32768),
                      ((40, 40, 40), 16384), ((56, 56, 40), 8192), ((40, 40, 42), 16384)):
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
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(stream)
        for _ in range(5): plan.compress(KEEP)
        e[1].record(stream)
        for _ in range(5): plan.decompress(odescs, pkg.WC_DEVICE)
        e[2].record(stream)
        torch.cuda.synchronize()
    K = plan.total_pairs()
    cms, dms = e[0].elapsed_time(e[1]) / 5, e[1].elapsed_time(e[2]) / 5
    gb = 8 * n * n_units / 1e9
    print(f"{dims} x {n_units}: {gb:.2f} GB f64, kept {K / (n * n_units):.3f}; compress {cms:.3f} ms = {gb / cms * 1e3:.0f} GB/s in, "
          f"alg {(8 * n * n_units + 8 * K) / cms / 1e6:.0f} GB/s; decompress {dms:.3f} ms, alg {(8 * K + 4 * n * n_units) / dms / 1e6:.0f} GB/s")
    plan.close(); del x, rec
