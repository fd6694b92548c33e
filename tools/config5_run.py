#!/usr/bin/env python3
"""BASELINE config 5: synthetic ~5 M-triangle height field, BVH in the BVH.py layout from the native builder,
4K frame.  Prints one JSON line per stage (build, upload, verify, render).
usage: config5_run.py [quads=1582] [width=3840] [height=2160] [spp=8] [verify_res=256]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from tests import fixtures  # noqa: E402
from tests.synthetic import height_field_scene  # noqa: E402

PARAMS = dict(cam_x="0", cam_y="-7.5", cam_z="4.5", cam_rx="-32", cam_ry="0", cam_rz="0", cam_DOF="50",
              sun_rx="60", sun_ry="0", sun_rz="30", sun_Power="0.8", IBL_Power="1.0")


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    a = [int(x) for x in sys.argv[1:]]
    quads, W, H, spp, vres = (a + [1582, 3840, 2160, 8, 256][len(a):])[:5]
    t0 = time.perf_counter()
    sc = height_field_scene(quads, seed=0)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    sc["BVH"], depth = rt.build_bvh(sc["faceData"], sc["V_p"], return_depth=True)
    t_bvh = time.perf_counter() - t0
    ntri = sc["faceData"].size // 10
    emit(stage="build", triangles=ntri, nodes=sc["BVH"].size // 9, depth=depth, generate_s=t_gen, bvh_build_s=t_bvh,
         threads=os.cpu_count())
    ctx = rt.Context(0)
    ibl = fixtures.load_ibl("preview")
    t0 = time.perf_counter()
    fixtures.upload(ctx, sc, ibl)
    emit(stage="upload", seconds=time.perf_counter() - t0)
    # correctness: fast traversal against the reference-order traversal, per ray, on a small frame
    cam, env = fixtures.cam_env(PARAMS, vres)
    o = rt.make_opts(rng_mode=rt.RNG_PHILOX, traversal=rt.TRAVERSAL_VERIFY, stack_cap=64, seed=0)
    ctx.render(cam, env, vres, vres, 2, 4, opts=o)
    st = ctx.stats()
    emit(stage="verify", res=vres, rays=st["rays"], mismatches=st["mismatches"])
    tri_f, k_f = ctx.primary_hits(cam, vres, vres, rt.make_opts(traversal=rt.TRAVERSAL_FAST))
    tri_r, k_r = ctx.primary_hits(cam, vres, vres, rt.make_opts(traversal=rt.TRAVERSAL_REFERENCE, stack_cap=64))
    emit(stage="primary", identical=bool(np.array_equal(tri_f, tri_r) and np.array_equal(k_f.view(np.uint32), k_r.view(np.uint32))),
         hit_frac=float((tri_f >= 0).mean()))
    cam, env = fixtures.cam_env(PARAMS, W, H)
    for rep in range(2):
        out = ctx.render(cam, env, W, H, spp, 4, opts=rt.make_opts(rng_mode=rt.RNG_PHILOX, seed=0, time_kernels=(rep == 1)))
        st = ctx.stats()
        emit(stage="render", rep=rep, width=W, height=H, spp=spp, rays=st["rays"], total_ms=st["total_ms"],
             primary_ms=st["primary_ms"], mrays_s=st["rays"] / st["total_ms"] / 1e3,
             msamples_s=W * H * spp / st["total_ms"] / 1e3, trace_kernel_ms=st["trace_kernel_ms"],
             shade_kernel_ms=st["shade_kernel_ms"], revalidated=st["revalidated"], mean=float(out.mean()),
             smem=st["scene_in_smem"])


if __name__ == "__main__":
    main()
