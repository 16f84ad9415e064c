import sys, time, numpy as np
sys.path.insert(0, '.')
import __graft_entry__ as g
import torch
pkg = g.package()
sys.argv += ['--x']
import bench
stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream=stream.cuda_stream)
tensors, descs, dims = bench.build_timestep_device(pkg, 0, torch.device('cuda', 0))
plan = ctx.plan(descs, pkg.WC_DEVICE)
outs = [torch.empty_like(tn, dtype=torch.float32) for tn in tensors]
optrs = []
for tn, lev in zip(outs, pkg.amr_synth.amr_levels()):
    n = lev.box ** 3
    optrs += [tn.data_ptr() + 4 * n * i for i in range(lev.n_boxes * bench.N_COMP)]
odescs = pkg.capi.box_descs(optrs, [pkg.WC_F32] * len(dims), dims)
with torch.cuda.stream(stream):
    plan.compress(bench.KEEP)
    for _ in range(2): plan.decompress(odescs, pkg.WC_DEVICE)
    torch.cuda.synchronize()
    ctx.set_profile(True); ctx.reset_counters()
    t0 = time.perf_counter()
    for _ in range(5): plan.decompress(odescs, pkg.WC_DEVICE)
    t1 = time.perf_counter()
    ctx.sync()
    t2 = time.perf_counter()
    print("host enqueue ms/step", (t1 - t0) * 200, "total ms/step", (t2 - t0) * 200)
    for k, (n, ms) in ctx.kernel_stats().items(): print(k, n, ms / 5)
    ctx.reset_counters()
    for _ in range(5): rm = plan.rmse(odescs)
    ctx.sync()
    for k, (n, ms) in ctx.kernel_stats().items(): print(k, n, ms / 5)
