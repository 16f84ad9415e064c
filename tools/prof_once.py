"""One warm pass of every hot kernel of the bench workload inside a cudaProfilerStart/Stop window (run under
`ncu --profile-from-start off`): compress (one timestep), the plan round-trip decompress, the stream decompress of the
`-d` path (second ctx, decode plan, segment index + decode kernels) and RMSE.  WC_OPT_PROFILE-free: ncu serialises the
kernels itself."""
import sys

import numpy as np

sys.path.insert(0, '.')
import __graft_entry__ as g
import torch

pkg = g.package()
capi = pkg.capi
import bench

stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream=stream.cuda_stream)
tensors, descs, dims = bench.build_timestep_device(pkg, 0, torch.device('cuda', 0))
torch.cuda.synchronize()
plan = ctx.plan(descs, pkg.WC_DEVICE)
ncoef = np.array([d[0] * d[1] * d[2] for d in dims], np.int64)
rec = torch.empty(int(ncoef.sum()), dtype=torch.float32, device='cuda')
rec2 = torch.empty(int(ncoef.sum()), dtype=torch.float32, device='cuda')
offs = np.concatenate([[0], np.cumsum(ncoef)])[:-1]
odescs = capi.box_descs([rec.data_ptr() + 4 * int(o) for o in offs], [pkg.WC_F32] * len(dims), dims)
sdescs = capi.box_descs([rec2.data_ptr() + 4 * int(o) for o in offs], [pkg.WC_F32] * len(dims), dims)
with torch.cuda.stream(stream):
    plan.compress(bench.KEEP)
hrec = plan.fetch_records(pkg.WC_HOST).copy()
k32 = hrec["npairs"].astype(np.int32)
total = int(k32.sum())
d_stream = torch.empty(max(total, 1), dtype=torch.int64, device='cuda')
capi.check(ctx.lib.wc_memcpy(ctx.h, d_stream.data_ptr(), int(hrec[0]["pairs"]), 8 * total, 0), "wc_memcpy", ctx.h)
d_k = torch.from_numpy(k32).cuda()
ctx2 = pkg.Context(0, stream=stream.cuda_stream)
dp = ctx2.decode_plan(sdescs, pkg.WC_DEVICE)
with torch.cuda.stream(stream):
    for _ in range(3):
        plan.compress(bench.KEEP)
        plan.decompress(odescs, pkg.WC_DEVICE)
        dp.decode(d_stream.data_ptr(), d_k.data_ptr(), pkg.WC_DEVICE)
    dp.finish()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    plan.compress(bench.KEEP)
    plan.decompress(odescs, pkg.WC_DEVICE)
    dp.decode(d_stream.data_ptr(), d_k.data_ptr(), pkg.WC_DEVICE)
    dp.finish()
    rm = plan.rmse(odescs)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok", float(rm.mean()), plan.total_pairs(), bool(torch.equal(rec, rec2)))
