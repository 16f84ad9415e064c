"""The `-d` path on the bench workload: dense device-resident pair stream -> wc_dplan_decode in a second ctx, per
unit class ('32', '64', 'all'), both segment-index kernels, per-kernel CUDA-event times; with a PHASE_PROFILE
build (WCGPU_LIB=... pointing at `make PHASE_PROFILE=1 OUT=...`) also the per-phase cycles of the decode kernels."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, '.')
import __graft_entry__ as g
import torch

pkg = g.package()
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
variants = [(0, int(c)) for c in sys.argv[2]] if len(sys.argv) > 2 else [(0, 1), (0, 0)]
sys.argv = sys.argv[:1]
import bench

capi = pkg.capi
stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream=stream.cuda_stream)
tensors, descs, dims = bench.build_timestep_device(pkg, 0, torch.device('cuda', 0))
torch.cuda.synchronize()
sel = [i for i, d in enumerate(dims) if which == 'all' or str(d[0]) == which]
plan = ctx.plan(descs[sel], pkg.WC_DEVICE)
sdims = [dims[i] for i in sel]
ncoef = np.array([d[0] * d[1] * d[2] for d in sdims], np.int64)
with torch.cuda.stream(stream):
    plan.compress(bench.KEEP)
hrec = plan.fetch_records(pkg.WC_HOST).copy()
k32 = hrec["npairs"].astype(np.int32)
total = int(k32.sum())
d_stream = torch.empty(max(total, 1), dtype=torch.int64, device='cuda')
capi.check(ctx.lib.wc_memcpy(ctx.h, d_stream.data_ptr(), int(hrec[0]["pairs"]), 8 * total, 0), "wc_memcpy", ctx.h)
d_k = torch.from_numpy(k32).cuda()
rec = torch.empty(int(ncoef.sum()), dtype=torch.float32, device='cuda')
offs = np.concatenate([[0], np.cumsum(ncoef)])
odescs = capi.box_descs([rec.data_ptr() + 4 * int(o) for o in offs[:-1]], [pkg.WC_F32] * len(sdims), sdims)
alg = int((8 * k32.astype(np.int64) + 4 * ncoef).sum())
for nn in sorted(set(ncoef.tolist())):
    kk = k32[ncoef == nn]
    print(f"units of {nn} cells: {kk.size}, K percentiles 1/10/50/90/99/max:", [int(x) for x in np.percentile(kk, [1, 10, 50, 90, 99, 100])])
peak = 6549.4
have_phase = hasattr(ctx.lib, "wc_debug_phase_cycles")
if have_phase:
    ctx.lib.wc_debug_phase_cycles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
out = (ctypes.c_ulonglong * 8)()
for seg, pipe in variants:
    ctx2 = pkg.Context(0, stream=stream.cuda_stream)
    ctx2.set_option(capi.WC_OPT_SEG_INDEX, seg)
    ctx2.set_option(capi.WC_OPT_DECODE_PIPE, pipe)
    dp = ctx2.decode_plan(odescs, pkg.WC_DEVICE)
    with torch.cuda.stream(stream):
        for _ in range(3):
            dp.decode(d_stream.data_ptr(), d_k.data_ptr(), pkg.WC_DEVICE)
        dp.finish()
        if have_phase:
            ctx.lib.wc_debug_phase_cycles(ctx2.h, out, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            dp.decode(d_stream.data_ptr(), d_k.data_ptr(), pkg.WC_DEVICE)
        e1.record(stream)
        dp.finish()
        ms = e0.elapsed_time(e1) / 10
        if have_phase:
            ctx.lib.wc_debug_phase_cycles(ctx2.h, out, 1)
        ctx2.set_profile(True)
        ctx2.reset_counters()
        for _ in range(5):
            dp.decode(d_stream.data_ptr(), d_k.data_ptr(), pkg.WC_DEVICE)
        dp.finish()
        st = ctx2.kernel_stats()
        ctx2.set_profile(False)
    print(f"{which} seg_index={seg} pipe={pipe}: {ms:.4f} ms/step  {alg / ms / 1e6:.0f} GB/s algorithmic = {alg / ms / 1e6 / peak:.3f} of {peak}",
          {k: round(v[1] / 5, 4) for k, v in st.items()})
    if have_phase:
        v = np.array(list(out), dtype=np.float64)
        units = max(v[5], 1)
        for n, c in zip(["top", "-", "decode (scan+scatter)", "barrier+stage/prefetch", "inverse+store"], v[:5]):
            print(f"    {n:22s} {c / units:9.0f} cycles/item  {100 * c / max(v[:5].sum(), 1):5.1f}%")
        print(f"    total cycles/item {v[:5].sum() / units:.0f}  (staged decode: waiting for the bulk copy {v[6] / units:.0f})  items/step {units / 10:.0f}")
    dp.close()
    ctx2.close()
