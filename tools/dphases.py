"""Phase cycle breakdown of the fused decompress kernels (needs a PHASE_PROFILE build: make -B PHASE_PROFILE=1)."""
import sys, ctypes, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as g
import torch
pkg = g.package()
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
sys.argv = sys.argv[:1]
import bench
stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream=stream.cuda_stream)
tensors, descs, dims = bench.build_timestep_device(pkg, 0, torch.device('cuda', 0))
torch.cuda.synchronize()
if which == '32': sel = [i for i, d in enumerate(dims) if d[0] == 32]
elif which == '64': sel = [i for i, d in enumerate(dims) if d[0] == 64]
else: sel = list(range(len(dims)))
plan = ctx.plan(descs[sel], pkg.WC_DEVICE)
sdims = [dims[i] for i in sel]
ncoef = [d[0] * d[1] * d[2] for d in sdims]
rec = torch.empty(sum(ncoef), dtype=torch.float32, device='cuda')
offs = np.concatenate([[0], np.cumsum(ncoef)])
odescs = pkg.capi.box_descs([rec.data_ptr() + 4 * int(o) for o in offs[:-1]], [pkg.WC_F32] * len(sdims), sdims)
lib = ctx.lib
lib.wc_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_ulonglong * 8)()
with torch.cuda.stream(stream):
    plan.compress(bench.KEEP)
    for _ in range(3): plan.decompress(odescs, pkg.WC_DEVICE)
    torch.cuda.synchronize()
    lib.wc_debug_phase_cycles(ctx.h, out, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): plan.decompress(odescs, pkg.WC_DEVICE)
    e1.record(stream)
    torch.cuda.synchronize()
    lib.wc_debug_phase_cycles(ctx.h, out, 1)
v = np.array(list(out), dtype=np.float64)
units = max(v[5], 1)
print(which, "decompress ms/step", e0.elapsed_time(e1) / 10, "cta-units", units / 10)
for n, c in zip(["zero-fill+barrier", "share sums (S>1)", "scan+scatter", "barrier+prefetch", "inverse+store"], v[:5]):
    print(f"  {n:20s} {c/units:9.0f} cycles/unit  {100*c/max(v[:5].sum(),1):5.1f}%")
print("  total cycles/unit", v[:5].sum() / units)
