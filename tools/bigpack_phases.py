"""Phase cycles of k_big_pack on 128^3 boxes (PHASE_PROFILE build: WCGPU_LIB=.../libwcgpu_prof.so)."""
import ctypes, sys
import numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as g
import torch
pkg = g.package()
stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream=stream.cuda_stream)
KEEP = 0.9990000128746033
dims, n_units = (128, 128, 128), 256
n = dims[0] * dims[1] * dims[2]
gen = torch.Generator(device='cuda'); gen.manual_seed(1)
x = torch.linspace(0, 50, n_units * n, device='cuda', dtype=torch.float64).sin_() * 100 + \
    torch.randn(n_units * n, device='cuda', dtype=torch.float64, generator=gen) * 0.05
descs = pkg.capi.box_descs([x.data_ptr() + 8 * n * i for i in range(n_units)], [pkg.WC_F64] * n_units, [dims] * n_units)
torch.cuda.synchronize()
plan = ctx.plan(descs, pkg.WC_DEVICE)
lib = ctx.lib
lib.wc_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_ulonglong * 8)()
with torch.cuda.stream(stream):
    plan.compress(KEEP)
    ctx.sync()
    lib.wc_debug_phase_cycles(ctx.h, out, 1)
    for _ in range(3): plan.compress(KEEP)
    ctx.sync()
    lib.wc_debug_phase_cycles(ctx.h, out, 1)
v = np.array(list(out), dtype=np.float64)
items = max(v[5], 1)
for nme, c in zip(["load (TMA)", "C1", "scan + look-back", "C2"], v[:4]):
    print(f"  {nme:18s} {c / items:9.0f} cycles/item")
print("  items", items / 3)
