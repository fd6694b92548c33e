#!/usr/bin/env python3
"""Headline benchmark of the path-tracing hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's own OpenCL-C kernels on the host cores

Workload = BASELINE.json configs[1]: Cornell box_Monkey, 1920x1080, 256 spp, maxBounce 4, glossy
monkey + glass wall + emitting lamp (the pinned material variant of SURVEY.md §8d), environment =
the only map in the reference checkout (600x300 preview; the 8k JPEG is missing upstream).
A "step" renders that whole frame once.  Scene buffers are the committed fixtures produced by the
reference's own FileManager/BVH.py (tests/golden/make_golden.py).

value      Mrays/s, device time (CUDA events on the launching stream), inputs resident in HBM.
e2e        the same metric through the reference-facing call with HOST buffers: every step re-uploads
           scene + environment (caches invalidated), renders, and reads the image back.
roofline   dominant kernel = k_trace (one launch per wavefront iteration).  HBM-style: algorithmic bytes of
           SURVEY.md §8d (36 B per box test + 136 B per triangle test of the REFERENCE traversal, + material
           bytes) of the rays the k_trace launches of one step process, over the summed duration of those
           launches (CUDA events around every launch, one extra profiled step), against the measured HBM peak.
           The scene is cache-resident by design, so this is a traffic-equivalent, not DRAM traffic
           (DESIGN.md §4); `traffic` is the ncu-measured DRAM bytes of one mid-frame k_trace launch.
cpu_baseline  oracle/_ref (the reference's .cl compiled by g++) on the host cores, bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tests import fixtures  # noqa: E402  (fixture loader only; no oracle import here)

WORKLOAD = dict(scene="monkey_cfg2", width=1920, height=1080, spp=256, max_bounce=4, ibl="preview", seed=0)
WORKLOAD_NAME = "Cornell box_Monkey 1920x1080 256spp maxBounce4 (glossy monkey, glass wall, emissive lamp)"


# --------------------------------------------------------------------------------------------------------
def load_workload():
    sc = fixtures.load_scene(WORKLOAD["scene"])
    ibl = fixtures.load_ibl(WORKLOAD["ibl"])
    cam, env = fixtures.cam_env(sc["params"], WORKLOAD["width"], WORKLOAD["height"])
    return sc, ibl, cam, env


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_renderer():
    """(kind, render(sc, cam, env, img_dim, spp, max_bounce, ibl, i0, i1) -> (image, rays)) of the CPU arm:
    oracle/_ref (the reference's own kernel text compiled by g++) when it is present, else the C restatement of
    oracle/ (bit-identical to it, tests/test_oracle_vs_ref.py).  The only place bench.py executes oracle/."""
    from oracle import ref_lib
    if ref_lib.available() and ref_lib.available("libclref_count.so"):
        def render(sc, cam, env, img_dim, spp, mb, ibl, i0, i1, count=False):
            out, cnt = ref_lib.raytrace(sc, cam, env, img_dim, spp, mb, ibl, i0=i0, i1=i1, counters=count)
            return out, (cnt["rays"] if cnt else None)
        return "reference", render
    from oracle import oracle

    def render(sc, cam, env, img_dim, spp, mb, ibl, i0, i1, count=False):
        out, cnt = oracle.render(sc, cam, env, img_dim, spp, mb, ibl, i0=i0, i1=i1)
        return out, cnt["rays"]
    return "port", render


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.thread = [], None, None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------------
def run_reference_arm(args, rank):
    """--impl reference: the reference's own Kernels/*.cl (compiled by g++ into oracle/_ref) on every
    host core, on a bounded sample of the workload per step.  Executes oracle/ — allowed only here."""
    if rank != 0:
        return
    kind, render = cpu_renderer()
    sc, ibl, cam, env = load_workload()
    W, H = WORKLOAD["width"], WORKLOAD["height"]
    cores = os.cpu_count() or 1
    spp = 16
    rows = max(8, min(H, 6 * cores))           # a band of rows through the middle of the frame
    i0 = (H // 2 - rows // 2) * W
    i1 = i0 + rows * W
    sample = f"{rows} rows x {W} px x {spp} spp of the workload frame (reference RNG), {cores} OpenMP threads"
    # ray count of the sample (untimed counting run; the kernel is deterministic)
    _, rays = render(sc, cam, env, W * H, spp, WORKLOAD["max_bounce"], ibl, i0, i1, count=True)
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        render(sc, cam, env, W * H, spp, WORKLOAD["max_bounce"], ibl, i0, i1)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    mrays = rays * len(times) / total / 1e6
    line = {"impl": "reference", "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_NAME, "sample": sample},
            "samples_per_s": (i1 - i0) * spp * len(times) / total,
            "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def cpu_baseline_leg(ctx, sc, ibl, cam, env):
    """Bounded CPU run of oracle/_ref beside the GPU number + parity of the same pixels."""
    import ensem3a_openclraytracer_b200 as rt
    kind, render = cpu_renderer()
    W, H = WORKLOAD["width"], WORKLOAD["height"]
    cores = os.cpu_count() or 1
    spp = 64
    rows = max(8, min(H, 6 * cores))
    i0 = (H // 2 - rows // 2) * W
    i1 = i0 + rows * W
    t0 = time.perf_counter()
    ref, _ = render(sc, cam, env, W * H, spp, WORKLOAD["max_bounce"], ibl, i0, i1)
    dt = time.perf_counter() - t0
    # same pixels, same (reference) generator on the GPU: ray count + parity
    o = rt.make_opts(rng_mode=rt.RNG_REFERENCE, pixel_begin=i0, pixel_end=i1)
    out = ctx.render(cam, env, W, H, spp, WORKLOAD["max_bounce"], opts=o)
    st = ctx.stats()
    a, b = out[3 * i0:3 * i1], ref[3 * i0:3 * i1]
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
    parity = {"pixels": i1 - i0, "spp": spp, "identical_frac": float(np.mean(a == b)), "max_rel": float(rel.max()),
              "rmse": float(np.sqrt(np.mean((a - b) ** 2)))}
    base = {"value": st["rays"] / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
            "sample": f"{rows} rows x {W} px x {spp} spp of the workload frame (reference RNG), {dt:.1f} s on {cores} "
                      f"OpenMP threads; samples/s {(i1 - i0) * spp / dt:.0f}"}
    return base, parity


# --------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200rt", choices=["b200rt", "reference"])
    ap.add_argument("--reduce", default="peer", choices=["nccl", "peer"], help="multi-GPU partial-sum reduction")
    ap.add_argument("--spp", type=int, default=WORKLOAD["spp"], help="debug only; the judged workload is 256")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    import ensem3a_openclraytracer_b200 as rt
    from ensem3a_openclraytracer_b200.multigpu import DistributedRenderer, split_range

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    sc, ibl, cam, env = load_workload()
    W, H, spp, mb = WORKLOAD["width"], WORKLOAD["height"], args.spp, WORKLOAD["max_bounce"]
    npix = W * H

    kl = rt.KernelLauncher(None, None, None, None, cuda_device=local)  # the reference-facing plugin object
    kl.rng_mode, kl.seed = rt.RNG_PHILOX, WORKLOAD["seed"]
    ctx = kl._ctx
    fixtures.upload(ctx, sc, ibl)
    stream = torch.cuda.Stream(device=dev)   # every launch, event and collective below is issued on this stream
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    out_dev = torch.zeros(npix * 3, dtype=torch.float32, device=dev)
    dr = DistributedRenderer(ctx, rank, world, reduce=args.reduce) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if dr is not None:
            dr.render(cam, env, W, H, spp, mb, rng_mode=rt.RNG_PHILOX, seed=WORKLOAD["seed"])
        else:
            ctx.render_device(cam, env, W, H, spp, mb, out_dev.data_ptr(),
                              rt.make_opts(rng_mode=rt.RNG_PHILOX, seed=WORKLOAD["seed"]))

    # ---- algorithmic bytes per ray: box / triangle tests of the REFERENCE traversal, counted on the GPU
    o = rt.make_opts(rng_mode=rt.RNG_PHILOX, traversal=rt.TRAVERSAL_REFERENCE, seed=WORKLOAD["seed"], collect_stats=True,
                     output=rt.OUT_SUMS, sample_begin=0, sample_end=2)
    ctx.render_device(cam, env, W, H, spp, mb, out_dev.data_ptr(), o)
    st = ctx.stats()
    box_per_ray, tri_per_ray = st["box_tests"] / st["rays"], st["tri_tests"] / st["rays"]
    hit_frac = 0.9  # upper bound on rays that fetch a 24 B material
    bytes_per_ray = 36.0 * box_per_ray + 136.0 * tri_per_ray + 24.0 * hit_frac

    # ---- device-resident timing ------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    rays_step = 0
    barrier()
    for a, b in ev:
        flush.fill_(1)                       # evict L2 between timed iterations (outside the event pair)
        a.record(stream)
        step_resident()
        b.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in ev)
    st = ctx.stats()
    rays_local = st["rays"]                 # last step, this rank
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    # every rank traces the frame's W*H primary rays (the cached primary hit); they count once for the job
    r = torch.tensor([float(rays_local - npix)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    ms_total, rays_step = float(t.item()), float(r.item()) + npix
    ms_per_step = ms_total / args.steps
    mrays = rays_step / ms_per_step / 1e3
    launches_step = st["kernel_launches"]

    # ---- one profiled step: CUDA events around every k_shade / k_trace launch (roofline of the dominant kernel) -----
    if dr is None:
        ctx.render_device(cam, env, W, H, spp, mb, out_dev.data_ptr(),
                          rt.make_opts(rng_mode=rt.RNG_PHILOX, seed=WORKLOAD["seed"], time_kernels=True))
    else:
        s0, s1 = split_range(spp, world)[rank]
        ctx.render_device(cam, env, W, H, spp, mb, out_dev.data_ptr(),
                          rt.make_opts(rng_mode=rt.RNG_PHILOX, seed=WORKLOAD["seed"], time_kernels=True,
                                       output=rt.OUT_SUMS, sample_begin=s0, sample_end=s1))
    barrier()
    stp = ctx.stats()

    # ---- end to end through the plugin with host buffers ---------------------------------------------------------
    host_out = np.zeros(npix * 3, np.float32)
    pinned = torch.empty(npix * 3, dtype=torch.float32).pin_memory() if world > 1 else None
    h2d = sum(int(sc[k].nbytes) for k in ("V_p", "V_n", "V_uv", "faceData", "materialData", "BVH")) + int(ibl.nbytes) + 60
    d2h = npix * 3 * 4

    def step_e2e():
        ctx.invalidate()
        if dr is None:
            kl.launch_Raytracing(host_out, sc["V_p"], sc["V_n"], sc["V_uv"], sc["faceData"], sc["materialData"],
                                 sc["lightData"], sc["BVH"], cam, env, npix, spp, mb, ibl)
        else:
            fixtures.upload(ctx, sc, ibl)
            ctx.set_stream(stream.cuda_stream)
            img = dr.render(cam, env, W, H, spp, mb, rng_mode=rt.RNG_PHILOX, seed=WORKLOAD["seed"])
            if rank == 0:
                pinned.copy_(img, non_blocking=True)
            torch.cuda.synchronize()

    ctx.set_stream(None if dr is None else stream.cuda_stream)
    e2e_steps = max(1, min(args.steps, 3))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_mrays = rays_step * e2e_steps / float(te.item()) / 1e6

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    n_trace = (launches_step - 1) // 2                      # k_primary + (n_iter + 1) k_shade + n_iter k_trace
    trace_rays = stp["rays"] - W * H                        # rank 0's rays minus the primary rays of k_primary
    achieved = trace_rays * bytes_per_ray / (stp["trace_kernel_ms"] * 1e-3) / 1e9
    traffic, ncu_view = None, None
    prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof):
        try:
            pj = json.load(open(prof))
            traffic = pj.get("k_trace_dram_bytes_per_launch")
            ncu_view = {k: pj[k] for k in ("k_trace_ipc_of_4", "k_trace_active_threads_per_instruction_of_32",
                                           "k_trace_l1tex_throughput_pct", "k_trace_alu_pipe_pct", "k_trace_fma_pipe_pct",
                                           "k_trace_l1_hit_pct", "binding_unit") if k in pj}
        except Exception:
            traffic = None
    line = {
        "metric": "Mrays/s", "value": mrays, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD_NAME, "width": W, "height": H, "spp": spp, "max_bounce": mb,
                   "rng": "philox4x32-10 keyed (pixel,sample,bounce)", "traversal": "fast",
                   "pipeline": "k_primary, then (spp*(maxBounce+2)) x (k_shade, k_trace) wavefront iterations",
                   "partition": "sample ranges" if world > 1 else "single GPU", "reduce": args.reduce if world > 1 else None,
                   "l2": "256 MiB flush write between timed iterations", "scene_in_smem": bool(st["scene_in_smem"])},
        "samples_per_s": npix * spp / (ms_per_step * 1e-3),
        "rays_per_step": rays_step, "rays_per_sample": rays_step / (npix * spp),
        "e2e": {"value": e2e_mrays, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * float(te.item()) / e2e_steps, "steps": e2e_steps},
        "gpu_launches": args.steps * (launches_step + (1 if world > 1 else 0)),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "bytes_per_ray": bytes_per_ray,
                     "box_tests_per_ray_reference": box_per_ray, "tri_tests_per_ray_reference": tri_per_ray,
                     "kernel": "k_trace", "launches_per_step": n_trace,
                     "avg_launch_us": 1e3 * stp["trace_kernel_ms"] / max(n_trace, 1),
                     "kernel_ms_per_step": stp["trace_kernel_ms"], "k_shade_ms_per_step": stp["shade_kernel_ms"],
                     "k_primary_ms_per_step": stp["primary_ms"], "step_ms_profiled": stp["total_ms"],
                     "share_of_step": stp["trace_kernel_ms"] / stp["total_ms"],
                     "note": "algorithmic bytes in the reference's layout and visiting order (SURVEY 8d); the 2 MB scene is "
                             "cache-resident, so this is a traffic-equivalent and frac may exceed 1; `ncu` holds the "
                             "committed profile of the same kernel (profiles/), which names the unit that binds",
                     "ncu": ncu_view},
    }
    if world == 1 and not args.no_cpu_baseline:
        ctx.set_stream(None)
        base, parity = cpu_baseline_leg(ctx, sc, ibl, cam, env)
        if base is not None:
            line["cpu_baseline"] = base
            line["parity_vs_reference"] = parity
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
