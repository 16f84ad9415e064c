"""Plan round trip (estimate mode) on the bench workload: per-kernel CUDA-event times of wc_plan_decompress for each
WC_OPT_DECODE_PIPE value given (default 0 1 2), then wc_plan_rmse."""
import sys, time, numpy as np
pipes = [int(a) for a in sys.argv[1:]] or [0, 1, 2]
sys.argv = sys.argv[:1]
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
    for pipe in pipes:
        ctx.set_option(pkg.capi.WC_OPT_DECODE_PIPE, pipe)
        for _ in range(2): plan.decompress(odescs, pkg.WC_DEVICE)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): plan.decompress(odescs, pkg.WC_DEVICE)
        e1.record(stream)
        torch.cuda.synchronize()
        ctx.set_profile(True); ctx.reset_counters()
        t0 = time.perf_counter()
        for _ in range(5): plan.decompress(odescs, pkg.WC_DEVICE)
        t1 = time.perf_counter()
        ctx.sync()
        print(f"pipe={pipe}: {e0.elapsed_time(e1) / 10:.4f} ms/step (host enqueue {(t1 - t0) * 200:.3f} ms/step)",
              {k: round(ms / 5, 4) for k, (n, ms) in ctx.kernel_stats().items()})
        ctx.set_profile(False)
    ctx.set_profile(True); ctx.reset_counters()
    for _ in range(5): rm = plan.rmse(odescs)
    ctx.sync()
    for k, (n, ms) in ctx.kernel_stats().items(): print(k, n, ms / 5)
