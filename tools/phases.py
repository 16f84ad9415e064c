import sys, ctypes, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as g
import torch
pkg = g.package()
sys.argv += ['--x']
import bench
stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream=stream.cuda_stream)
tensors, descs, dims = bench.build_timestep_device(pkg, 0, torch.device('cuda', 0))
torch.cuda.synchronize()
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
if which == '32': sel = [i for i, d in enumerate(dims) if d[0] == 32]
elif which == '64': sel = [i for i, d in enumerate(dims) if d[0] == 64]
else: sel = list(range(len(dims)))
if len(sys.argv) > 2 and sys.argv[2].isdigit(): sel = sel[:int(sys.argv[2])]   # few units: the input stays in L2 (hot-input diagnostic)
plan = ctx.plan(descs[sel], pkg.WC_DEVICE)
lib = ctx.lib
have_prof = hasattr(lib, 'wc_debug_phase_cycles')   # only in `make prof` builds; the product library just gets timed
if have_prof: lib.wc_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_ulonglong * 8)()
with torch.cuda.stream(stream):
    for _ in range(3): plan.compress(bench.KEEP)
    torch.cuda.synchronize()
    if have_prof: lib.wc_debug_phase_cycles(ctx.h, out, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): plan.compress(bench.KEEP)
    e1.record(stream)
    torch.cuda.synchronize()
    if have_prof: lib.wc_debug_phase_cycles(ctx.h, out, 1)
if not have_prof:
    print(which, 'ms/step', e0.elapsed_time(e1) / 10)
    sys.exit(0)
v = np.array(list(out), dtype=np.float64)
units = v[5]
print(which, "ms/step", e0.elapsed_time(e1)/10, "units", units/10)
names = ["A(transform+wait)", "B(threshold)", "C1(count)", "scan/exchange", "C2(emit)"]
for n, c in zip(names, v[:5]):
    print(f"  {n:20s} {c/units:9.0f} cycles/unit  {100*c/v[:5].sum():5.1f}%")
print("  total cycles/unit", v[:5].sum()/units)
print(f"  cluster waits (inside B / scan): {v[6]/units:.0f} / {v[7]/units:.0f} cycles/unit")
