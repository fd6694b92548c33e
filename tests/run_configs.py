#!/usr/bin/env python3
"""TEST-SIDE MEASUREMENT SCRIPT (it executes oracle/_ref as the CPU arm and parity checker, which only code
under tests/ and bench.py may do).  BASELINE.json configs 1-4 at their named sizes on one GPU (config 5:
tools/config5_run.py; config 2 is also bench.py's workload).  One JSON line per config: device time, Mrays/s, Msamples/s, kernel split, and a bounded
CPU run of the reference's own kernels (oracle/_ref, all host threads) on a band of the same frame with parity.
usage: python tests/run_configs.py [spp_scale=1.0]   (spp_scale < 1 shortens configs 3 and 4)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from oracle import ref_lib  # noqa: E402   (checker / CPU baseline only)
from tests import fixtures  # noqa: E402


CONFIGS = [
    dict(id=1, scene="cornell", w=512, h=512, spp=16, rng=rt.RNG_REFERENCE, ibl="preview", note="reference RNG (seed = pixel index)"),
    dict(id=1, scene="cornell", w=512, h=512, spp=16, rng=rt.RNG_PHILOX, ibl="preview", note="Philox"),
    dict(id=2, scene="monkey_cfg2", w=1920, h=1080, spp=256, rng=rt.RNG_PHILOX, ibl="preview", note="bench workload"),
    dict(id=3, scene="furnace_cfg3", w=1920, h=1080, spp=1024, rng=rt.RNG_PHILOX, ibl="grey", note="uniform grey IBL, sun 0", scale=True),
    dict(id=4, scene="serre", w=3840, h=2160, spp=512, rng=rt.RNG_PHILOX, ibl="8k", note="8192x4096 stand-in IBL", scale=True),
    dict(id="1080p target", scene="cornell", w=1920, h=1080, spp=256, rng=rt.RNG_PHILOX, ibl="preview", note="north_star: >= 1 Gray/s on the Cornell box at 1080p"),
]


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    ctx = rt.Context(0)
    cores = os.cpu_count() or 1
    for c in CONFIGS:
        sc = fixtures.load_scene(c["scene"])
        ibl = fixtures.load_ibl(c["ibl"])
        fixtures.upload(ctx, sc, ibl)
        W, H = c["w"], c["h"]
        spp = max(4, int(round(c["spp"] * scale))) if c.get("scale") else c["spp"]
        cam, env = fixtures.cam_env(sc["params"], W, H)
        best = None
        for rep in range(2):
            out = ctx.render(cam, env, W, H, spp, 4, opts=rt.make_opts(rng_mode=c["rng"], seed=0, time_kernels=(rep == 1)))
            st = ctx.stats()
            if best is None or st["total_ms"] < best["total_ms"]:
                best = dict(st)
            if rep == 1:
                split = (st["trace_kernel_ms"], st["shade_kernel_ms"], st["primary_ms"])
        line = dict(config=c["id"], scene=c["scene"], width=W, height=H, spp=spp, note=c["note"], rays=best["rays"],
                    device_ms=best["total_ms"], mrays_s=best["rays"] / best["total_ms"] / 1e3,
                    msamples_s=W * H * spp / best["total_ms"] / 1e3, rays_per_sample=best["rays"] / (W * H * spp),
                    k_trace_ms=split[0], k_shade_ms=split[1], k_primary_ms=split[2], revalidated=best["revalidated"],
                    image_mean=float(out.mean()))
        # bounded CPU run of the reference kernels on a band of rows + parity of those pixels (reference RNG)
        if ref_lib.available():
            rows = max(4, min(H, 2 * cores))
            cspp = min(spp, 8)
            i0 = (H // 2 - rows // 2) * W
            i1 = i0 + rows * W
            t0 = time.perf_counter()
            ref, _ = ref_lib.raytrace(sc, cam, env, W * H, cspp, 4, ibl, i0=i0, i1=i1)
            dt = time.perf_counter() - t0
            got = ctx.render(cam, env, W, H, cspp, 4, opts=rt.make_opts(rng_mode=rt.RNG_REFERENCE, pixel_begin=i0, pixel_end=i1))
            st = ctx.stats()
            a, b = got[3 * i0:3 * i1], ref[3 * i0:3 * i1]
            rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
            line.update(cpu_mrays_s=st["rays"] / dt / 1e6, cpu_threads=cores, cpu_sample=f"{rows} rows x {W} px x {cspp} spp",
                        parity_identical_frac=float(np.mean(a == b)), parity_max_rel=float(rel.max()))
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
