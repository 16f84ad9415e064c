"""One warm compress + decompress + rmse of the bench workload inside a cudaProfilerStart/Stop window
(run under `ncu --profile-from-start off`).  WC_OPT_PROFILE-free: ncu serialises the kernels itself."""
import sys
sys.path.insert(0, '.')
import __graft_entry__ as g
import torch
pkg = g.package()
import bench
stream = torch.cuda.Stream()
ctx = pkg.Context(0, stream=stream.cuda_stream)
tensors, descs, dims = bench.build_timestep_device(pkg, 0, torch.device('cuda', 0))
torch.cuda.synchronize()
plan = ctx.plan(descs, pkg.WC_DEVICE)
ncoef = [d[0] * d[1] * d[2] for d in dims]
rec = torch.empty(sum(ncoef), dtype=torch.float32, device='cuda')
offs = [0]
for n in ncoef: offs.append(offs[-1] + n)
odescs = pkg.capi.box_descs([rec.data_ptr() + 4 * o for o in offs[:-1]], [pkg.WC_F32] * len(dims), dims)
with torch.cuda.stream(stream):
    for _ in range(3):
        plan.compress(bench.KEEP)
        plan.decompress(odescs, pkg.WC_DEVICE)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    plan.compress(bench.KEEP)
    plan.decompress(odescs, pkg.WC_DEVICE)
    rm = plan.rmse(odescs)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok", float(rm.mean()), plan.total_pairs())
