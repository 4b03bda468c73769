#!/usr/bin/env python
"""Benchmark of the hot path: full DM_saliency pipeline (grid build + normals/curvature + dk/dn +
normalise/blend) on a synthetic 10 M-point terrain at 3 neighbourhood radii (BASELINE.json configs[1]).

    python bench.py --gpus 1 --steps 5 --warmup 3            # our arm (one JSON line on stdout)
    python bench.py --impl reference --gpus 1 ...            # reference CPU path on the host cores
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # weak scaling

A "step" = one pass of the whole pipeline over the cloud (all 3 radii).  `value` = points/s with the
cloud resident in HBM; `e2e` = the same through the public tile API (`multiscale_saliency_tiles`) with
HOST buffers: every step's cloud is copied H2D from pinned memory and its three saliency columns D2H
inside the timed region, the copies of neighbouring steps overlapping the kernels.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RADII = (0.5, 0.75, 1.0)
PARAMS = dict(roughness=0.02, alpha=0.05, sigma_normal=0.01, sigma_curvature=0.01, min_points=4)
CW = 0.5
METRIC = "saliency points/sec (10M pts, 3 radii)"
UNIT = "points/s"


def workload_name(n):
    return "configs[1]: full DM_saliency (curvature + dk/dn + normalise/blend) on %d synthetic terrain points, radii %s" % (
        n, "/".join(str(r) for r in RADII))


def config_dict(n, world):
    """The `config` object of the JSON line -- the SAME for our arm and the reference arm."""
    return {"workload": workload_name(n), "points_per_gpu": n, "radii": list(RADII),
            "l2": "inputs (240 MB cloud + 320..640 MB records per radius) exceed the 126 MB L2",
            "parallelism": "x-strips with 2*r_max ghost halos over NCCL" if world > 1 else "single GPU"}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.idx = gpu_index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nme, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        os.unlink(self.f.name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path on all host cores: the UNMODIFIED reference from
    oracle/_ref (sklearn BallTree.query_radius + verbatim Curvature_Kernel / Saliency_Kernel callbacks) when that
    was built, else the oracle port.  Each step is a bounded sample of Q query points per radius against the
    full cloud's tree, extrapolated x N/Q (Q = 50 000)."""
    if rank != 0:
        return
    from more3d_b200 import synthetic
    from oracle.cpu_bench import CpuSaliencyBaseline, host_cores, describe
    n = args.n
    pts = synthetic.terrain(n, seed=2)
    cores = host_cores()
    base = CpuSaliencyBaseline(pts, RADII, PARAMS, workers=cores, seed=1)
    # Q query points per radius and step (50 000); with many steps the per-step sample shrinks so
    # that the whole run stays within a few minutes (>= 300 000 sampled queries per radius in total)
    q = args.cpu_q if args.steps <= 6 else max(10000, min(args.cpu_q, 300000 // args.steps))
    for _ in range(args.warmup):
        base.step(max(q // 8, cores))
    times = []
    for _ in range(args.steps):
        total, raw = base.step(q)
        times.append(total)
    base.close()
    t = float(np.mean(times))
    value = n / t
    sample = ("%s; tree over all %d points built once (%.1f s, included in every step), %d random query points per radius "
              "per step on %d worker processes, extrapolated x N/Q = %.0f" % (describe(base.kind), n, base.t_tree, q, cores,
                                                                           n / float(q)))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(n, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": base.kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from more3d_b200 import synthetic, _lib
    from more3d_b200.pipeline import multiscale_saliency, multiscale_saliency_tiles

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    n = args.n
    _lib.load()

    if distributed:
        from more3d_b200.distributed import StripSaliency
        job = StripSaliency(n, rank, world, dev, RADII, PARAMS, CW, seed=2)
        step_resident = job.step_resident
        steps_e2e = job.steps_e2e
        h2d, d2h = job.h2d_bytes, job.d2h_bytes
        sum_m = job.sum_m
    else:
        pts = synthetic.terrain(n, seed=2)
        host = torch.from_numpy(pts).pin_memory()
        resident = host.to(dev)
        out_host = {r: torch.empty(n, dtype=torch.float32).pin_memory() for r in RADII}
        h2d = host.numel() * 8
        d2h = sum(t.numel() * 4 for t in out_host.values())

        def step_resident():
            return multiscale_saliency(resident, RADII, curvature_weight=CW, **PARAMS)

        def steps_e2e(k):
            # the public tile API: k host-resident tiles one after the other; every tile is copied H2D from pinned
            # memory and its three saliency columns D2H inside the timed region, the copies of neighbouring tiles
            # overlapping the kernels (the same synthetic tile is sent k times)
            return multiscale_saliency_tiles([host] * k, RADII, out_hosts=[out_host] * k, curvature_weight=CW,
                                             keep_results=False, **PARAMS)

        # neighbour counts for the algorithmic-bytes figure (outside the timed region)
        res = multiscale_saliency(resident, RADII, curvature_weight=CW, keep=True, **PARAMS)
        sum_m = {r: float(res[r]["_pcount"].double().sum()) for r in RADII}
        nonzero = {r: (float((res[r]["_dk"] != 0).double().mean()), float((res[r]["_dn"] != 0).double().mean()))
                   for r in RADII}
        gpu_normals = res[RADII[0]]["normals"].cpu().numpy() if args.cpu_baseline else None
        gpu_K = res[RADII[0]]["_nonparam_K"].cpu().numpy() if args.cpu_baseline else None
        del res

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if distributed:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    _lib.profile_reset()
    _lib.profile_enable(True)
    launches0 = _lib.launch_count()
    if rank == 0:
        sampler.start()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - launches0
    prof = _lib.profile_all()
    _lib.profile_enable(False)
    def timed_batch(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(steps)                      # returns with every copy of every step completed
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if distributed:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # what this box's host <-> device path delivers from the SAME pinned buffers (outside the timed region): the e2e leg
    # overlaps the copies with the kernels, so a step costs max(kernels, copies) and a slow path shows up here first
    # (with several GPUs every rank probes at the same time, as in the e2e leg; rank 0 reports what it saw)
    def copy_rate(dst, src, nbytes):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dst.copy_(src, non_blocking=True)
        barrier()
        a.record()
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        return 3 * nbytes / (a.elapsed_time(b) * 1e-3) / 1e9
    p_host = job.host if distributed else host
    p_out = job.out_host if distributed else out_host
    tmp = torch.empty((n, 3), dtype=torch.float64, device=dev)
    col = torch.zeros(n, dtype=torch.float32, device=dev)
    copy_probe = {"h2d_GBs": copy_rate(tmp, p_host, h2d), "d2h_GBs": copy_rate(p_out[RADII[0]], col, 4 * n)}
    del tmp, col

    # warm-up of the e2e path with >= 3 tiles: two tiles are in flight in the stream, so the caching allocator (blocks
    # handed to the copy streams come back late) and the pinned status buffers reach their steady state only after a few
    ms_e2e_warm = timed_batch(steps_e2e, max(3, min(args.warmup, 5)))
    ms_e2e = timed_batch(steps_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    del ms_e2e_warm

    total_pts = float(n) * world
    value = total_pts * args.steps / (ms * 1e-3)
    e2e_value = total_pts * args.steps / (ms_e2e * 1e-3)

    # roofline of the dominant kernel (algorithmic bytes: SURVEY.md §8d)
    peak, peak_src = measured_peak()
    alg = {"dkdn": sum(n * 40.0 + 32.0 * sum_m[r] for r in RADII),
           "moments": sum(n * 48.0 + 16.0 * sum_m[r] for r in RADII)}
    kern = {}
    for name, (tot_ms, cnt) in prof.items():
        kern[name] = {"ms_per_step": tot_ms / args.steps, "launches_per_step": cnt / args.steps}
        if name in alg and tot_ms > 0:
            kern[name]["algorithmic_GBs"] = alg[name] * args.steps / (tot_ms * 1e-3) / 1e9
    dom = max((k for k in kern if k in alg), key=lambda k: kern[k]["ms_per_step"])
    achieved = kern[dom]["algorithmic_GBs"]
    tr = ncu_traffic()
    names = {"dkdn": "dk/dn pass: dkdn_screen_kernel + dk_refine_kernel + dkdn_refine_kernel", "moments": "moments_kernel"}
    calls = len(RADII)
    traffic = (tr or {}).get(dom)
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src,
                "traffic": traffic, "algorithmic_bytes_per_step": alg[dom], "calls_per_step": calls,
                "kernel_share_of_step": kern[dom]["ms_per_step"] / (ms / args.steps),
                # what the memory system really moves: ncu DRAM bytes per call (profiles/traffic.json) over the live time
                "dram_GBs": (traffic * calls / (kern[dom]["ms_per_step"] * 1e-3) / 1e9) if traffic else None,
                # the same figure for both hot groups (moments = one kernel, three launches)
                "per_group": {k: {"achieved": kern[k]["algorithmic_GBs"], "frac": kern[k]["algorithmic_GBs"] / peak,
                                  "ms_per_step": kern[k]["ms_per_step"], "algorithmic_bytes_per_step": alg[k]}
                              for k in kern if k in alg and "algorithmic_GBs" in kern[k]},
                "note": ("achieved = SURVEY 8d pair-gather bytes (N*40 + 32*sum m per radius) / CUDA-event time of the group; "
                         "the gather is served by L1/L2 (DRAM traffic = `traffic` per call), so this HBM-equivalent rate can "
                         "exceed the HBM peak: the binding limits are issue slots and the FP64 pipe (profiles/README.md)")}

    # ---- configs[4]: strong scaling of ONE 500 M-point tile (r = 0.5), the same global cloud for every N ----
    strong = None
    if args.strong_points > 0:
        if distributed:
            del job
        else:
            del resident, host, out_host
        import gc
        from more3d_b200.pipeline import release_staging_buffers
        release_staging_buffers()
        gc.collect()
        torch.cuda.empty_cache()
        strong = strong_leg(args, rank, world, dev)

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(n, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps, "copy_probe": copy_probe},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": kern}
    if not distributed:
        line["nonzero_fraction_dk_dn"] = {str(r): nonzero[r] for r in RADII}
    if strong is not None:
        line["strong"] = strong

    if args.cpu_baseline and not distributed:
        # the reference's CPU path on the host cores beside the GPU number (same run, same cloud): the UNMODIFIED
        # reference when oracle/_ref was built, else the oracle port; bounded sample, extrapolated x N/Q
        from oracle.cpu_bench import CpuSaliencyBaseline, host_cores, describe
        cores = host_cores()
        base = CpuSaliencyBaseline(pts, RADII, PARAMS, workers=cores, normals=gpu_normals, K=gpu_K, seed=1)
        total, raw = base.step(args.cpu_q)
        base.close()
        line["cpu_baseline"] = {
            "value": n / total, "unit": UNIT, "cores": cores, "kind": base.kind,
            "sample": ("%s; tree over all %d points (%.1f s) + %d random query points per radius (%.1f s on %d worker "
                       "processes), extrapolated x N/Q = %.0f" % (describe(base.kind), n, base.t_tree, args.cpu_q, raw, cores,
                                                                  n / float(args.cpu_q)))}
        del base
    else:
        line["cpu_baseline"] = None

    if args.configs and not distributed:
        # BASELINE.json configs[0], [2], [3] at full size on this GPU, each with its algorithmic-bytes fraction, a
        # bounded CPU leg and a parity spot-check (tools/bench_configs.py)
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        try:
            del pts, gpu_normals, gpu_K
        except NameError:
            pass
        line["configs"] = {}
        for key in ("0", "2", "3"):
            try:
                line["configs"][key] = bench_configs.run_all((key,), small=args.small_configs, peak=peak)[key]
            except Exception as e:  # a failing sub-record must not lose the headline line
                line["configs"][key] = {"error": "%s: %s" % (type(e).__name__, e)}
            torch.cuda.empty_cache()
    print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


def strong_leg(args, rank, world, dev):
    """BASELINE.json configs[4]: the saliency pipeline at r = 0.5 on ONE tile of `--strong-points` points, cut
    into `world` x-strips (radius-halo exchange + min/max + histogram all-reduce over NCCL).  The tile is the
    counter-based synthetic terrain (point i is a function of (seed, i), generated in HBM), so every N processes
    the SAME global cloud: `checksum` (sum of neighbour counts, sum of the saliency bit patterns) must not depend
    on N.  Timing as the main line: CUDA events, barrier + synchronize on both sides, max over ranks."""
    import torch
    import torch.distributed as dist
    from more3d_b200 import synthetic
    from more3d_b200.pipeline import multiscale_saliency
    total, r = int(args.strong_points), 0.5
    if total % world or synthetic.HASH_SLABS % world:
        return {"skipped": "strong_points / slab count not divisible by the number of GPUs"}
    per = total // world
    lx, ly = synthetic.hash_tile_extent(total)
    own = synthetic.hash_terrain_device(5, rank * per, per, total, device=dev)
    distributed = world > 1
    if distributed:
        from more3d_b200.distributed import StripSaliency
        x_lo = rank * lx / world - 0.5 * lx
        job = StripSaliency(per, rank, world, dev, (r,), PARAMS, CW, cloud=own, x_range=(x_lo, x_lo + lx / world),
                            host_buffers=False)
        step = job.step_resident
        cs = torch.tensor(list(job.checksum[r]), dtype=torch.int64, device=dev)
        dist.all_reduce(cs)
    else:
        res = multiscale_saliency(own, (r,), curvature_weight=CW, keep=True, **PARAMS)
        cs = torch.stack([res[r]["_pcount"].long().sum(), res[r]["_saliency"].view(torch.int32).long().sum()])
        del res

        def step():
            return multiscale_saliency(own, (r,), curvature_weight=CW, **PARAMS)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = max(min(args.steps, 3), 1)
    a.record()
    for _ in range(k):
        step()
    b.record()
    barrier()
    ms = a.elapsed_time(b)
    if distributed:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    cs = [int(v) for v in cs.cpu().tolist()]
    out = {"workload": "configs[4]: %d-point synthetic tile, DM_saliency pipeline at r=0.5, x-strips over %d GPU(s)" % (total, world),
           "points": total, "n_gpus": world, "scaling": "strong", "steps": k, "ms_per_step": ms / k,
           "points_per_s": total * k / (ms * 1e-3),
           "checksum": {"sum_pcount": cs[0], "sum_saliency_f32_bits_mod_2^64": cs[1]},
           "hbm_allocated_GB_rank0": torch.cuda.max_memory_allocated() / 1e9}
    del own
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", dest="n", type=int, default=10_000_000, help="points per GPU")
    ap.add_argument("--cpu-q", type=int, default=50000,
                    help="CPU legs: query points per radius per step, split over the host cores (BASELINE.md section 3: 50 000)")
    ap.add_argument("--strong-points", type=int, default=500_000_000,
                    help="size of the configs[4] strong-scaling tile (0 = skip the leg)")
    ap.add_argument("--no-configs", dest="configs", action="store_false",
                    help="skip the configs[0]/[2]/[3] sub-records (N = 1 only)")
    ap.add_argument("--small-configs", action="store_true", help="sub-records at reduced sizes (smoke runs)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--radii", default=None, help="comma-separated radii (default 0.5,0.75,1.0 = configs[1]); "
                                                  "'--points 62500000 --radii 0.5' on 8 GPUs is configs[4] (500 M-point tile)")
    args = ap.parse_args()
    if args.radii:
        global RADII
        RADII = tuple(float(x) for x in args.radii.split(","))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        if world == 1 and args.gpus > 1:
            # launched without torchrun: re-exec under it so `python bench.py --gpus N` works too
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
            raise SystemExit(subprocess.call(cmd))
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
