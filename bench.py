#!/usr/bin/env python
"""bench.py — compress / decompress throughput of the B200 numeric core on the AMR-256-L4 workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path (forward Haar + threshold + (run,value) packing,
src/compressor.cpp:203-247) over one synthetic AMReX timestep: 256^3 level 0 (64 boxes of 64^3) +
3 refinement levels (512 boxes of 32^3 each), 8 float64 components = 12 800 units, 4.29 GB of input
field data, keep = 0.999f (BASELINE.json configs[2]).  Every rank keeps a short SERIES of distinct
timesteps resident (BASELINE configs[3]: the timestep-sharded series; rank r owns t = r*T .. r*T+T-1) and
steps through it with wc_plan_set_inputs + wc_plan_compress on ONE plan; there is no data-path collective
(units are independent) and scaling is weak.

value    : GB/s of float64 input field data, whole job, inputs resident in HBM, outputs left in HBM.
e2e      : the same metric through the C-ABI plan with HOST (pinned) input boxes and the packed stream fetched
           back to pinned host memory: H2D + kernels + dense gather + D2H inside the timed region.  Also
           e2e.f32 (host boxes already float32, as the reference's multiBox3D is, src/preprocess.cpp:78) and
           e2e.h2d_ceiling_gbs: the same call with WC_OPT_COPY_ONLY (same pinned buffers, same chunking, no
           kernels) = what the host link allows.
roofline : the dominant kernel's algorithmic bytes (8N + 8K + 20 per unit) / its average launch
           duration measured with CUDA events on the launch stream (WC_OPT_PROFILE) against
           MEASURED_PEAKS.json's hbm_gbs.
decompress_stream : the `-d` path (src/decompressor.cpp:238-255): a SECOND ctx decodes the dense, device-resident
           pair stream through a decode plan (wc_dplan_*) — no compress-side segment tables, no cached decode
           tables; the segment-index kernel is inside the timed region.  With its own roofline (8K + 4N
           algorithmic bytes) and e2e (host pairs -> host float32 boxes).
decompress_roundtrip : wc_plan_decompress on the compressing plan (estimate mode: tables from the compress kernels).
cpu_baseline / --impl reference: the reference's own compress() (oracle/_ref, unmodified sources) on the host
           cores over a bounded sample of the same units.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KEEP = float(np.float32(0.999))
N_COMP = 8
UNITS_PER_TIMESTEP = 12800
WORKLOAD = ("AMR-256-L4: 256^3 level 0 (64 boxes of 64^3) + 3 refinement levels (512 boxes of 32^3 each), "
            "8 float64 components, 12800 units, 4.29 GB per timestep, keep=0.999f")


def make_config(args):
    """The static description of the workload — identical for both arms (the driver compares it)."""
    return {"workload": WORKLOAD, "keep": KEEP, "units_per_gpu": UNITS_PER_TIMESTEP,
            "series": f"{args.timesteps} distinct timesteps resident per rank (rank r: t = r*T .. r*T+T-1), one per step, "
                      "through wc_plan_set_inputs on one plan",
            "l2": "inputs (4.29 GB per step) larger than L2; no flush needed", "path": args.path}


def load_pkg():
    import __graft_entry__ as g
    return g.package()


# ------------------------------------------------------------------------------------------------
# clocks sampling (NVML in a thread; nvidia-smi has too coarse a period for millisecond steps)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), threading.Event()
        self.ok, self.max_mhz = False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2, "display_clock_setting": 0x100}
        while not self.stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.ok:
            self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.ok:
            self.th.join()

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def build_timestep_device(pkg, t, device):
    """One AMR-256-L4 timestep generated on the device, FAB order.  Returns (tensors, descs, dims)."""
    synth = pkg.amr_synth
    levels = synth.amr_levels()
    tensors, ptrs, dims = [], [], []
    for lev in levels:
        fab = synth.generate_level_torch(lev, N_COMP, t=t, device=device)  # (boxes, comp, z, y, x) f64
        tensors.append(fab)
        n = lev.box ** 3
        base = fab.data_ptr()
        for b in range(lev.n_boxes):
            for c in range(N_COMP):
                ptrs.append(base + 8 * n * (b * N_COMP + c))
                dims.append((lev.box,) * 3)
    descs = pkg.capi.box_descs(ptrs, [pkg.WC_F64] * len(ptrs), dims)
    return tensors, descs, dims


def cpu_sample_units(pkg, t=0, l0_boxes=8, fine_boxes=32):
    """Bounded sample of the same workload on the host: (float32 box, dims) per unit."""
    synth = pkg.amr_synth
    units = []
    for lev in synth.amr_levels():
        fab = synth.generate_level_numpy(lev, N_COMP, t=t)
        nb = l0_boxes if lev.level == 0 else fine_boxes
        step = max(1, lev.n_boxes // nb)
        for b in list(range(0, lev.n_boxes, step))[:nb]:
            for c in range(N_COMP):
                units.append((fab[b, c].astype(np.float32), (lev.box,) * 3))  # src/preprocess.cpp:78
        del fab
    return units


SAMPLE_DESC = "8 of 64 level-0 boxes + 32 of 512 boxes per fine level of timestep 0, 8 components"


def time_reference_cpu(units, steps, warmup, stub_lzma, threads):
    """Reference compress() (oracle/_ref when present, else the C port) over `units`, `threads`
    host threads over disjoint units.  Returns (seconds per step list, kind)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import pyoracle
    kind = "reference" if pyoracle.have_ref() else "port"
    if kind == "reference":
        ref = pyoracle.Ref()
        ref.set_lzma_mode(1 if stub_lzma else 0)
    else:
        orc = pyoracle.Oracle()
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    times = []
    with tempfile.TemporaryDirectory(dir=shm) as d, ThreadPoolExecutor(threads) as pool:
        def work(i):
            box, dims = units[i]
            if kind == "reference":
                ref.compress(box.reshape(1, -1), dims, KEEP, d, t=0, lev=0, box_idx=i, want_pairs=False)
            else:
                orc.compress_unit(box, dims, KEEP)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            list(pool.map(work, range(len(units)), chunksize=max(1, len(units) // (threads * 8))))
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    if kind == "reference":
        ref.set_lzma_mode(0)
    return times, kind


def normalise_steps(args):
    """Both arms run exactly these counts (W >= 3 is the timing rule; K >= 1)."""
    return max(args.warmup, 3), max(args.steps, 1)


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores, honouring
    --steps / --warmup; each step is a bounded sample of the workload (about 0.1-0.2 s of CPU work)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pkg = load_pkg()
    threads = os.cpu_count() or 1
    units = cpu_sample_units(pkg)
    field_bytes = sum(8 * b.size for b, _ in units)  # float64 field data the units came from
    W, K = normalise_steps(args)
    times, kind = time_reference_cpu(units, K, W, stub_lzma=True, threads=threads)
    sec = statistics.median(times)
    value = field_bytes / sec / 1e9
    sub = units[::16]
    full_times, _ = time_reference_cpu(sub, 1, 0, stub_lzma=False, threads=threads)
    full_value = sum(8 * b.size for b, _ in sub) / full_times[0] / 1e9
    sample = (f"{len(units)} units ({SAMPLE_DESC}; {field_bytes / 1e6:.0f} MB of f64 field data) per step; reference "
              f"compress() (src/compressor.cpp:192-297) with the LZMA stage stubbed out and files on tmpfs = numeric "
              f"core + serialisation only; with LZMA (xz preset 6) it drops to {full_value:.4f} GB/s")
    line = {"impl": "reference", "metric": "compress GB/s of input field data", "value": value, "unit": "GB/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(args),
            "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": kind, "sample": sample,
                             "with_lzma_value": full_value},
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "e2e_with_lzma": {"value": full_value, "unit": "GB/s",
                              "sample": f"{len(sub)} of the sampled units, full compress() incl. xz preset 6"},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# main arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 generic kernels only, 2 fused only")
    ap.add_argument("--timesteps", type=int, default=4, help="distinct timesteps resident per rank (the series)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-lzma", action="store_true")
    ap.add_argument("--no-irregular", action="store_true", help="skip the irregular-box leg (x-slab kernels)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--seg-index", type=int, default=0, help="WC_OPT_SEG_INDEX for the stream decompress leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    pkg = load_pkg()
    capi = pkg.capi
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libwcgpu has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    affinity = pin_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    W, K = normalise_steps(args)
    T = max(1, args.timesteps)

    stream = torch.cuda.Stream(device=device)
    ctx = pkg.Context(local, stream=stream.cuda_stream)
    ctx.set_path(args.path)
    lib = ctx.lib
    series = [build_timestep_device(pkg, t=rank * T + i, device=device) for i in range(T)]
    torch.cuda.synchronize()
    tensors, descs, dims = series[0]
    n_units = len(descs)
    ncoef = np.array([d[0] * d[1] * d[2] for d in dims], np.int64)
    field_bytes = int(8 * ncoef.sum())
    plan = ctx.plan(descs, pkg.WC_DEVICE)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step(i):
        plan.set_inputs(series[i % T][1])
        plan.compress(KEEP)

    # per-timestep pair counts (untimed): the algorithmic bytes of a step depend on K
    npairs_t = []
    for i in range(T):
        step(i)
        npairs_t.append(plan.fetch_records(pkg.WC_DEVICE)["npairs"].astype(np.int64).copy())

    # ---- timed region: device resident, one distinct timestep per step ---------------------------------
    with torch.cuda.stream(stream):
        for i in range(W):
            step(i)
        barrier()
        ctx.reset_counters()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            barrier()
            ev0.record(stream)
            for i in range(K):
                step(i)
            ev1.record(stream)
            barrier()
        ms_total = ev0.elapsed_time(ev1)
        launches = ctx.counter(capi.WC_CTR_KERNEL_LAUNCHES)
    ms_per_step = max_over_ranks(ms_total) / K
    value = world * field_bytes / (ms_per_step * 1e-3) / 1e9
    alg_steps = [int((8 * ncoef + 8 * npairs_t[i % T] + 20).sum()) for i in range(K)]
    alg_all = sum(alg_steps) / K
    kept_fraction = float(np.mean([npairs_t[i].sum() / ncoef.sum() for i in range(T)]))

    # ---- roofline: per-kernel device time with CUDA events on the launch stream ------------------------
    ctx.set_profile(True)
    ctx.reset_counters()
    prof_steps = 2 * T
    for i in range(prof_steps):
        step(i)
    ctx.sync()
    stats = ctx.kernel_stats()
    ctx.set_profile(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak, peak_src = (peaks["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs") if "hbm_gbs" in peaks else (6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)")
    stats.pop("k_patch_inputs", None)
    step_kernel_ms = sum(ms for _, ms in stats.values()) / prof_steps
    dom = max(stats.items(), key=lambda kv: kv[1][1]) if stats else ("none", (1, 0.0))
    dom_name, (dom_n, dom_ms) = dom
    is32, is64 = ncoef == 32 ** 3, ncoef == 64 ** 3
    np_mean = np.mean(np.stack(npairs_t), axis=0)                 # profiled steps visit every timestep equally
    alg_by_kernel = {
        "k_fused_compress<1,cube32>": float((8 * ncoef[is32] + 8 * np_mean[is32] + 20).sum()),
        "k_fused_compress<8,cube64>": float((8 * ncoef[is64] + 8 * np_mean[is64] + 20).sum()),
        "k_forward_generic": float((8 * ncoef).sum()),          # reads the f64 input once (writes 4N scratch)
        "k_emit_tiles": float((8 * np_mean).sum()),
        "k_count_tiles": 0.0,
    }
    dom_alg = alg_by_kernel.get(dom_name, alg_all)
    dom_avg_ms = dom_ms / max(dom_n, 1)
    dom_alg_per_launch = dom_alg / max(dom_n / prof_steps, 1)
    achieved = dom_alg_per_launch / (dom_avg_ms * 1e-3) / 1e9 if dom_avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "alg_bytes_per_launch": dom_alg_per_launch, "avg_launch_ms": dom_avg_ms,
                "kernel_share_of_step": (dom_ms / prof_steps) / step_kernel_ms if step_kernel_ms else None,
                "whole_step": {"alg_bytes": alg_all, "achieved": alg_all / (ms_per_step * 1e-3) / 1e9,
                               "frac": alg_all / (ms_per_step * 1e-3) / 1e9 / peak},
                "kernels_ms_per_step": {k: ms / prof_steps for k, (n, ms) in stats.items()},
                "kernels_frac_alone": {k: alg_by_kernel[k] / ((ms / prof_steps) * 1e-3) / 1e9 / peak
                                       for k, (n, ms) in stats.items() if alg_by_kernel.get(k) and ms > 0}}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get(dom_name)
        except Exception:  # noqa: BLE001
            pass

    # ---- from here on: timestep 0 of this rank --------------------------------------------------------
    plan.set_inputs(descs)
    plan.compress(KEEP)
    rec = plan.fetch_records(pkg.WC_DEVICE).copy()
    npairs = rec["npairs"].astype(np.int64)
    total_pairs = int(npairs.sum())
    outs = [torch.empty_like(tn, dtype=torch.float32) for tn in tensors]
    optrs = []
    for tn, lev in zip(outs, pkg.amr_synth.amr_levels()):
        n = lev.box ** 3
        optrs += [tn.data_ptr() + 4 * n * i for i in range(lev.n_boxes * N_COMP)]
    odescs = capi.box_descs(optrs, [pkg.WC_F32] * n_units, dims)
    dec_alg = int((8 * npairs + 4 * ncoef).sum())

    def timed(fn, warm, steps):
        with torch.cuda.stream(stream):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(steps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / steps

    dsteps = max(3, min(K, 20))
    # (1) estimate-mode round trip: the decoder reads the segment tables the compress kernels wrote
    rt_ms = max_over_ranks(timed(lambda: plan.decompress(odescs, pkg.WC_DEVICE), 2, dsteps))
    decompress_roundtrip = {"ms_per_step": rt_ms, "alg_bytes": dec_alg, "achieved": dec_alg / (rt_ms * 1e-3) / 1e9,
                            "frac": dec_alg / (rt_ms * 1e-3) / 1e9 / peak,
                            "note": "wc_plan_decompress on the compressing plan (estimate mode): segment tables come "
                                    "from the compress kernels, device tables cached"}
    rm = plan.rmse(odescs)
    rmse_ms = timed(lambda: plan.rmse(odescs), 0, 3)
    rmse_alg = int((12 * ncoef).sum())          # float64 original + float32 reconstruction, read once
    rmse_info = {"ms_per_step": rmse_ms, "alg_bytes": rmse_alg, "achieved": rmse_alg / (rmse_ms * 1e-3) / 1e9,
                 "frac": rmse_alg / (rmse_ms * 1e-3) / 1e9 / peak,
                 "note": "wc_plan_rmse incl. the D2H of the per-unit results"}

    # (2) the real `-d` path: a DIFFERENT ctx decodes the dense device-resident stream through a decode plan
    hrec = plan.fetch_records(pkg.WC_HOST).copy()                       # dense pinned host stream, unit after unit
    h_pairs_addr = int(hrec[0]["pairs"])
    d_stream = torch.empty(max(total_pairs, 1), dtype=torch.int64, device=device)
    capi.check(lib.wc_memcpy(ctx.h, d_stream.data_ptr(), h_pairs_addr, 8 * total_pairs, 0), "wc_memcpy", ctx.h)
    k32 = npairs.astype(np.int32)
    d_k = torch.from_numpy(k32).to(device)
    souts = [torch.empty_like(tn, dtype=torch.float32) for tn in tensors]
    sptrs = []
    for tn, lev in zip(souts, pkg.amr_synth.amr_levels()):
        n = lev.box ** 3
        sptrs += [tn.data_ptr() + 4 * n * i for i in range(lev.n_boxes * N_COMP)]
    sdescs = capi.box_descs(sptrs, [pkg.WC_F32] * n_units, dims)
    ctx2 = pkg.Context(local, stream=stream.cuda_stream)
    ctx2.set_path(args.path)
    ctx2.set_option(capi.WC_OPT_SEG_INDEX, args.seg_index)
    dplan = ctx2.decode_plan(sdescs, pkg.WC_DEVICE)

    def stream_decode():
        dplan.decode(d_stream.data_ptr(), d_k.data_ptr(), pkg.WC_DEVICE)
    sd_ms = max_over_ranks(timed(stream_decode, 3, dsteps))
    dplan.finish()
    ctx2.set_profile(True)
    ctx2.reset_counters()
    for _ in range(5):
        stream_decode()
    dplan.finish()
    sstats = ctx2.kernel_stats()
    ctx2.set_profile(False)
    s_launches = sum(n for n, _ in sstats.values()) / 5
    decompress_stream = {
        "ms_per_step": sd_ms, "value": world * (4 * int(ncoef.sum())) / (sd_ms * 1e-3) / 1e9,
        "unit": "GB/s of float32 output field data", "alg_bytes": dec_alg,
        "api": "wc_dplan_decode (second ctx, dense device-resident pair stream + counts, no tables, no cache)",
        "roofline": {"bound": "hbm", "achieved": dec_alg / (sd_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": dec_alg / (sd_ms * 1e-3) / 1e9 / peak,
                     "frac_of_nominal_8tbs": dec_alg / (sd_ms * 1e-3) / 1e9 / 8000.0,
                     "alg": "8K + 4N per unit; traffic of the 64^3-class units is 16K + 4N (the index kernel reads "
                            "their lists once more)"},
        "kernels_ms_per_step": {k: ms / 5 for k, (n, ms) in sstats.items()},
        "launches_per_step": s_launches, "seg_index": args.seg_index}

    # ---- parity spot check of this very run against the oracle (not timed) -----------------------------
    parity = None
    cpu_baseline = None
    if rank == 0:
        from oracle.pyoracle import Oracle
        orc = Oracle()
        checked, ok, ok_stream, worst = 0, True, True, 0.0
        flat_in = [tn.reshape(-1) for tn in tensors]
        unit_level = np.repeat(np.arange(4), [64 * N_COMP] + [512 * N_COMP] * 3)
        first = np.concatenate([[0], np.cumsum([64 * N_COMP] + [512 * N_COMP] * 3)])
        for u in list(range(0, n_units, 997)) + [n_units - 1]:
            lv = int(unit_level[u])
            n = int(ncoef[u])
            off = (u - int(first[lv])) * n
            box = flat_in[lv][off:off + n].cpu().numpy()
            runs, vals, _ = orc.compress_unit(box, dims[u], KEEP)
            k = int(rec[u]["npairs"])
            got = np.empty(max(k, 1), capi.PAIR)
            if k:
                capi.check(lib.wc_memcpy(ctx.h, got.ctypes.data, int(rec[u]["pairs"]), 8 * k, 1), "wc_memcpy", ctx.h)
            ok &= (k == runs.size and got["run"][:k].tobytes() == runs.tobytes() and got["val"][:k].tobytes() == vals.tobytes())
            ob = orc.decompress_unit(runs, vals, dims[u])
            rec_box = outs[lv].reshape(-1)[off:off + n].cpu().numpy()
            ok &= rec_box.tobytes() == ob.reshape(-1).tobytes()
            ok_stream &= souts[lv].reshape(-1)[off:off + n].cpu().numpy().tobytes() == ob.reshape(-1).tobytes()
            oe = orc.rmse(box.astype(np.float32), ob, dims[u])
            worst = max(worst, abs(rm[u] - oe) / max(abs(oe), 1e-300))
            checked += 1
        parity = {"units_checked": checked, "pairs_and_recon_bit_exact": bool(ok),
                  "stream_decompress_bit_exact": bool(ok_stream), "rmse_max_rel_err": worst,
                  "rmse_tolerance": 1e-12, "mean_rmse_per_component": [float(np.mean(rm[c::N_COMP])) for c in range(N_COMP)]}
    if world > 1:
        gt_ok, gq_ok = global_threshold_parity(pkg, ctx, stream, device, rank, world)
        if rank == 0:
            parity["global_threshold_ok"] = gt_ok
            parity["global_quantile_ok"] = gq_ok
            parity["global_quantile_note"] = ("quantile extension: radix-select histograms of the sharded batch summed over "
                                              "NCCL, every rank's pairs == oracle threshold_pack at the numpy quantile")
            parity["global_threshold_note"] = ("config-5 extension: one threshold for a seeded batch sharded over the ranks, "
                                               "NCCL MAX all-reduce of the arg-max key, every rank's pairs == oracle "
                                               "threshold_pack with the oracle's concatenation-rule threshold")

    # ---- e2e: host boxes in, packed stream out, through the C ABI -------------------------------------
    e2e = None
    e2e_dec = None
    if not args.no_e2e:
        es = max(1, args.e2e_steps)

        def host_copy(dtype_code):
            """Pinned host copy of timestep 0 (same values), float64 or float32 (src/preprocess.cpp:78)."""
            esz = 8 if dtype_code == pkg.WC_F64 else 4
            hp = ctypes.c_void_p()
            capi.check(lib.wc_host_alloc(ctypes.byref(hp), esz * int(ncoef.sum())), "wc_host_alloc")
            off = 0
            for tn in tensors:
                src = tn if esz == 8 else tn.to(torch.float32)
                torch.cuda.synchronize()
                nb = src.numel() * esz
                capi.check(lib.wc_memcpy(ctx.h, hp.value + off, src.data_ptr(), nb, 1), "wc_memcpy", ctx.h)
                off += nb
                del src
            ptrs, off = [], 0
            for d in dims:
                ptrs.append(hp.value + off)
                off += esz * d[0] * d[1] * d[2]
            return hp, capi.box_descs(ptrs, [dtype_code] * n_units, dims)

        def time_host_plan(hplan):
            barrier()
            ctx.reset_counters()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(es):
                hr = hplan.compress_to_host_records(KEEP)
            e1.record(stream)
            barrier()
            ms = max(e0.elapsed_time(e1) / es, (time.perf_counter() - t0) * 1e3 / es)
            return max_over_ranks(ms), ctx.counter(capi.WC_CTR_H2D_BYTES) // es, ctx.counter(capi.WC_CTR_D2H_BYTES) // es, hr

        hp64, hdescs = host_copy(pkg.WC_F64)
        hplan = ctx.plan(hdescs, pkg.WC_HOST)
        hplan.compress_to_host_records(KEEP)  # warm-up (allocates the dense buffers)
        e_ms, h2d, d2h, hr = time_host_plan(hplan)
        same = bool(np.array_equal(hr["npairs"], rec["npairs"]))
        # copy-only probe: same pinned buffers, same chunking, kernels skipped -> the host-link ceiling
        ctx.set_option(capi.WC_OPT_COPY_ONLY, 1)
        c_ms, _, _, _ = time_host_plan(hplan)
        ctx.set_option(capi.WC_OPT_COPY_ONLY, 0)
        e2e = {"value": world * field_bytes / (e_ms * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": es,
               "api": "wc_plan_compress_to_host (pipelined H2D / kernels / gather + D2H) on pinned float64 host boxes",
               "same_pair_counts_as_device_run": same,
               "h2d_ceiling_gbs": world * field_bytes / (c_ms * 1e-3) / 1e9, "copy_only_ms_per_step": c_ms,
               "frac_of_ceiling": c_ms / e_ms,
               "ceiling_note": "WC_OPT_COPY_ONLY: the same call, same pinned buffers and chunking, no kernels",
               "cpu_affinity": affinity}
        hplan.close()
        lib.wc_host_free(hp64)
        # float32 host boxes: what the reference's host actually holds (multiBox3D, src/preprocess.cpp:78)
        hp32, hdescs32 = host_copy(pkg.WC_F32)
        hplan32 = ctx.plan(hdescs32, pkg.WC_HOST)
        hplan32.compress_to_host_records(KEEP)
        f_ms, fh2d, fd2h, hr32 = time_host_plan(hplan32)
        e2e["f32"] = {"value": world * field_bytes / (f_ms * 1e-3) / 1e9, "unit": "GB/s of float64-equivalent field data",
                      "value_f32_bytes": world * (field_bytes // 2) / (f_ms * 1e-3) / 1e9, "ms_per_step": f_ms,
                      "h2d_bytes_per_step": int(fh2d), "d2h_bytes_per_step": int(fd2h),
                      "same_pair_counts_as_device_run": bool(np.array_equal(hr32["npairs"], rec["npairs"])),
                      "note": "host boxes already narrowed to float32 (the reference's multiBox3D): half the H2D bytes"}
        hplan32.close()
        lib.wc_host_free(hp32)
        # decompress e2e: host pair stream (pinned) -> host float32 boxes (pinned)
        hb = ctypes.c_void_p()
        capi.check(lib.wc_host_alloc(ctypes.byref(hb), 4 * int(ncoef.sum())), "wc_host_alloc")
        hptrs, off = [], 0
        for d in dims:
            hptrs.append(hb.value + off)
            off += 4 * d[0] * d[1] * d[2]
        hod = capi.box_descs(hptrs, [pkg.WC_F32] * n_units, dims)
        hdplan = ctx2.decode_plan(hod, pkg.WC_HOST)
        hdplan.decode(h_pairs_addr, k32.ctypes.data, pkg.WC_HOST)
        hdplan.finish()
        barrier()
        ctx2.reset_counters()
        t0 = time.perf_counter()
        for _ in range(es):
            hdplan.decode(h_pairs_addr, k32.ctypes.data, pkg.WC_HOST)
            hdplan.finish()
        barrier()
        d_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / es)
        chk = np.frombuffer((ctypes.c_char * (4 * int(ncoef[0]))).from_address(hptrs[0]), np.float32)
        e2e_dec = {"value": world * (4 * int(ncoef.sum())) / (d_ms * 1e-3) / 1e9, "unit": "GB/s of float32 output field data",
                   "ms_per_step": d_ms, "h2d_bytes_per_step": int(ctx2.counter(capi.WC_CTR_H2D_BYTES) // es),
                   "d2h_bytes_per_step": int(ctx2.counter(capi.WC_CTR_D2H_BYTES) // es), "steps": es,
                   "api": "wc_dplan_decode + wc_dplan_finish: pinned host pair stream -> pinned host float32 boxes "
                          "(H2D pairs | kernels | D2H boxes pipelined over 8 chunks)",
                   "first_box_matches_device_run": bool(chk.tobytes() == souts[0].reshape(-1)[:int(ncoef[0])].cpu().numpy().tobytes())}
        decompress_stream["e2e"] = e2e_dec
        hdplan.close()
        lib.wc_host_free(hb)

    # ---- e2e incl. the host LZMA stage, overlapped with the GPU chunks (bounded sample, rank 0) ----------
    e2e_lzma = None
    if rank == 0 and world == 1 and not args.no_lzma and not args.no_e2e:
        e2e_lzma = lzma_leg(pkg, ctx)

    # ---- boxes the 16-byte vector kernels refuse (nz % 4 != 0, odd dimensions): the x-slab kernels ---------
    irregular = None
    if rank == 0 and world == 1 and not args.no_irregular:
        try:
            irregular = irregular_leg(pkg, local, stream, device, peak, args.path)
        except Exception as e:      # a side leg must not take the headline line down with it
            irregular = {"error": f"{type(e).__name__}: {e}"}

    # ---- CPU baseline: the reference's own code on this box's host cores ---------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        units = cpu_sample_units(pkg)
        fb = sum(8 * b.size for b, _ in units)
        times, kind = time_reference_cpu(units, 2, 1, stub_lzma=True, threads=threads)
        sec = min(times)
        sub = units[::16]
        full_times, _ = time_reference_cpu(sub, 1, 0, stub_lzma=False, threads=threads)
        cpu_baseline = {"value": fb / sec / 1e9, "unit": "GB/s", "cores": threads, "kind": kind,
                        "sample": (f"{len(units)} units ({SAMPLE_DESC}, {fb / 1e6:.0f} MB f64 field data), reference compress() "
                                   f"with LZMA stubbed + tmpfs files (numeric core + serialisation)"),
                        "with_lzma_value": sum(8 * b.size for b, _ in sub) / full_times[0] / 1e9,
                        "with_lzma_sample": f"{len(sub)} of those units, full compress() incl. xz preset 6"}

    if rank == 0:
        cfg = make_config(args)
        line = {"metric": "compress GB/s of input field data", "value": value, "unit": "GB/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "kept_fraction": kept_fraction,
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu_baseline,
                "decompress_stream": decompress_stream, "decompress_roundtrip": decompress_roundtrip,
                "rmse": rmse_info, "e2e_with_lzma": e2e_lzma, "irregular_boxes": irregular, "parity": parity}
        print(json.dumps(line))
    dplan.close()
    ctx2.close()
    plan.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def irregular_leg(pkg, local, stream, device, peak, path):
    """Off-BASELINE shapes, reported beside the headline: boxes with nz % 4 != 0 or odd dimensions (the trailing element
    passes through, src/compressor.cpp:98-175, and comes back as 0, src/decompressor.cpp:99-108) on the x-slab kernels of
    csrc/wc_xslab.cu — device-resident compress and plan round-trip decompress per shape, algorithmic bytes 8N + 8K and
    8K + 4N, and a bit-for-bit check of a few units against the oracle (not timed)."""
    import torch
    capi = pkg.capi
    from oracle.pyoracle import Oracle
    orc = Oracle()
    ctx = pkg.Context(local, stream=stream.cuda_stream)
    ctx.set_path(path)
    out = {}

    def timed(fn, warm=2, steps=5):
        with torch.cuda.stream(stream):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(steps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / steps

    for dims in ((40, 40, 42), (31, 17, 9), (65, 61, 57)):
        n = dims[0] * dims[1] * dims[2]
        n_units = max(1, (1 << 29) // (8 * n))                      # ~0.5 GB of float64 input per shape (> L2)
        gen = torch.Generator(device=device)
        gen.manual_seed(17)
        x = torch.linspace(0, 50, n_units * n, device=device, dtype=torch.float64).sin_() * 100 + \
            torch.randn(n_units * n, device=device, dtype=torch.float64, generator=gen) * 0.05
        rec = torch.empty(n_units * n, dtype=torch.float32, device=device)
        descs = capi.box_descs([x.data_ptr() + 8 * n * i for i in range(n_units)], [pkg.WC_F64] * n_units, [dims] * n_units)
        odescs = capi.box_descs([rec.data_ptr() + 4 * n * i for i in range(n_units)], [pkg.WC_F32] * n_units, [dims] * n_units)
        torch.cuda.synchronize()
        plan = ctx.plan(descs, pkg.WC_DEVICE)
        c_ms = timed(lambda: plan.compress(KEEP))
        k = plan.total_pairs()
        d_ms = timed(lambda: plan.decompress(odescs, pkg.WC_DEVICE))
        ctx.sync()
        plan.close()
        # parity of the first units (a small plan of its own: fetching every unit's pairs would dominate the leg)
        m = min(3, n_units)
        small = ctx.plan(descs[:m], pkg.WC_DEVICE)
        small.compress(KEEP)
        small.decompress(odescs[:m], pkg.WC_DEVICE)
        ctx.sync()
        got = small.fetch_host()
        ok = True
        for i in range(m):
            b = x[i * n:(i + 1) * n].cpu().numpy().reshape(dims[2], dims[1], dims[0])
            runs, vals, _ = orc.compress_unit(b, dims, KEEP)
            ok = ok and got[i].runs.tobytes() == runs.tobytes() and got[i].vals.tobytes() == vals.tobytes()
            ob = orc.decompress_unit(runs, vals, dims)
            ok = ok and rec[i * n:(i + 1) * n].cpu().numpy().tobytes() == ob.tobytes()
        small.close()
        c_alg, d_alg = 8 * n * n_units + 8 * k, 8 * k + 4 * n * n_units
        out["x".join(map(str, dims))] = {
            "units": n_units, "kept_fraction": k / (n * n_units),
            "compress_ms": c_ms, "compress_alg_gbs": c_alg / (c_ms * 1e-3) / 1e9, "compress_frac": c_alg / (c_ms * 1e-3) / 1e9 / peak,
            "decompress_ms": d_ms, "decompress_alg_gbs": d_alg / (d_ms * 1e-3) / 1e9,
            "decompress_frac": d_alg / (d_ms * 1e-3) / 1e9 / peak, "bit_exact_vs_oracle": bool(ok)}
        del x, rec
    ctx.close()
    out["note"] = ("plan compress / plan round-trip decompress of ~0.5 GB of float64 boxes per shape, CUDA events; fractions of "
                   "the measured HBM peak on algorithmic bytes (8N + 8K, 8K + 4N)")
    return out


def pin_to_gpu_numa_node(index):
    """Bind this rank's threads to the CPUs NVML reports as closest to its GPU BEFORE any pinned allocation, so that the
    e2e legs' pinned buffers come from the GPU's own NUMA node (first-touch) and the copies do not cross sockets."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return {"cpus": len(os.sched_getaffinity(0)), "how": "nvmlDeviceSetCpuAffinity"}
    except Exception as e:  # noqa: BLE001
        return {"cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None, "how": "unchanged: " + repr(e)[:80]}


def lzma_leg(pkg, ctx):
    """compress() end to end INCLUDING the host LZMA stage (src/compressor.cpp:256-291) on the bounded sample the CPU
    arm uses: the chunk callback of wc_plan_compress_to_host_chunked hands finished units to a host thread pool while
    later chunks are still on the GPU."""
    import lzma
    from concurrent.futures import ThreadPoolExecutor
    capi = pkg.capi
    units = cpu_sample_units(pkg)
    boxes = [b for b, _ in units]
    dims = [d for _, d in units]
    fb = sum(8 * b.size for b in boxes)
    threads = os.cpu_count() or 1
    hplan = ctx.plan_host(boxes, dims)
    hplan.compress_to_host_records(KEEP)
    sizes = [0] * len(units)

    def encode(i, shape, ncoef, k, addr):
        head = np.array([*shape, ncoef, k], dtype="<i4").tobytes()
        body = ctypes.string_at(addr, 8 * k) if k else b""
        sizes[i] = len(lzma.compress(head + body, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC64, preset=6))

    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as pool:
        futs = []

        def on_chunk(first, n, recs):
            for j in range(n):
                r = recs[j]
                futs.append(pool.submit(encode, first + j, [int(s) for s in r["shape"]], int(r["ncoef"]), int(r["npairs"]),
                                        int(r["pairs"])))
        hplan.compress_to_host_chunked(KEEP, on_chunk)
        t_gpu = time.perf_counter() - t0
        for f in futs:
            f.result()
    sec = time.perf_counter() - t0
    hplan.close()
    return {"value": fb / sec / 1e9, "unit": "GB/s", "seconds": sec, "gpu_call_seconds": t_gpu, "host_threads": threads,
            "xz_bytes": int(sum(sizes)), "field_bytes": int(fb),
            "sample": f"{len(units)} units ({SAMPLE_DESC}; float32 host boxes), xz preset 6 / CRC64 per unit on {threads} "
                      f"host threads fed by the chunk callback while the GPU works on later chunks",
            "note": "LZMA-bound: the numeric core is no longer visible in this figure on either arm"}


def global_threshold_parity(pkg, ctx, stream, device, rank, world):
    """Config-5 extension across ranks, once, untimed: a seeded batch (identical on every rank) is sharded, each rank
    transforms its units, the arg-max keys are all-reduced over NCCL and every rank packs with the winner.  Checked
    against the oracle's concatenation rule (threshold) and threshold_pack (pairs) on every rank; AND-reduced."""
    import torch
    import torch.distributed as dist
    from oracle.pyoracle import Oracle
    orc = Oracle()
    rng = np.random.default_rng(515)
    dims = [(64, 64, 64)] * 4 + [(32, 32, 32)] * 28
    boxes = []
    for i, d in enumerate(dims):
        X, Y, Z = d
        ii, jj, kk = np.meshgrid(np.arange(X), np.arange(Y), np.arange(Z), indexing="ij")
        f = 50 * np.sin(0.1 * ii + 0.3 * i) * np.cos(0.07 * jj) * np.sin(0.05 * kk + 0.1) + 1e-3 * rng.standard_normal(ii.shape)
        boxes.append(np.ascontiguousarray(f.transpose(2, 1, 0)) * (1.0 + 0.37 * (i % 5)) * (-1.0 if i % 2 else 1.0))
    lo, hi = pkg.amr_synth.shard_units([b.size for b in boxes], world, rank)
    mine = [torch.from_numpy(b).to(device) for b in boxes[lo:hi]]
    descs = pkg.capi.box_descs([t.data_ptr() for t in mine], [pkg.WC_F64] * len(mine), dims[lo:hi])
    keep = float(np.float32(0.99))
    ok = True
    with torch.cuda.stream(stream):
        plan = ctx.plan(descs, pkg.WC_DEVICE)
        pkg.distributed.compress_global_threshold(plan, keep, lo, device)
        got = plan.fetch_host()
        plan.close()
        # the quantile extension: radix-select histograms summed over NCCL (plan with a scratch for every unit)
        ctx.set_path(1)
        plan = ctx.plan(descs, pkg.WC_DEVICE)
        pkg.distributed.compress_global_quantile(plan, keep, sum(b.size for b in boxes[lo:hi]), device)
        got_q = plan.fetch_host()
        plan.close()
        ctx.set_path(0)
    flats = [orc.haar_forward(orc.narrow(b), d) for b, d in zip(boxes, dims)]
    tg = orc.select_threshold_global(flats, keep)
    for i, p in enumerate(got):
        rg, vg = orc.threshold_pack(flats[lo + i], tg)
        ok &= p.runs.tobytes() == rg.tobytes() and p.vals.tobytes() == vg.tobytes()
    from oracle.pyoracle import quantile_threshold
    tq = quantile_threshold(flats, keep)
    okq = True
    for i, p in enumerate(got_q):
        rq, vq = orc.threshold_pack(flats[lo + i], tq)
        okq &= p.runs.tobytes() == rq.tobytes() and p.vals.tobytes() == vq.tobytes()
    flag = torch.tensor([1 if ok else 0, 1 if okq else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return bool(flag[0].item()), bool(flag[1].item())


def _run():
    # Libraries (NCCL's version banner, torchrun) write to fd 1; the contract is ONE JSON line on stdout.
    # Everything except our own final print goes to stderr.
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real, "w")
    global print
    import builtins

    def json_print(*a, **k):
        builtins.print(*a, **k, file=out)
        out.flush()
    print = json_print
    return main()


if __name__ == "__main__":
    sys.exit(_run())
