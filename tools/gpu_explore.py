#!/usr/bin/env python3
"""First-contact exploration on a B200: parity against the oracle + timings, printed as JSON lines.
(Development tool; the judged checks are tests/ -m gpu and bench.py.)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from oracle import oracle  # noqa: E402
from tests import fixtures  # noqa: E402


def emit(**kw):
    print(json.dumps(kw), flush=True)


def parity(ctx, name, res, spp, bounce, ibl_name="preview", rng=0, seed=0):
    sc = fixtures.load_scene(name)
    ibl = fixtures.load_ibl(ibl_name)
    fixtures.upload(ctx, sc, ibl)
    cam, env = fixtures.cam_env(sc["params"], res)
    n = res * res
    t0 = time.time()
    ref, cnt = oracle.render(sc, cam, env, n, spp, bounce, ibl, rng_mode=rng, seed=seed)
    t_cpu = time.time() - t0
    prim = oracle.primary(sc, cam, n)
    for trav, tname in ((0, "fast"), (1, "reference"), (2, "verify")):
        o = rt.make_opts(rng_mode=rng, traversal=trav, seed=seed, collect_stats=(trav != 2))
        tri, k = ctx.primary_hits(cam, res, res, o)
        out = ctx.render(cam, env, res, res, spp, bounce, opts=o)
        st = ctx.stats()
        rel = np.abs(out - ref) / np.maximum(np.abs(ref), 1e-3)
        emit(kind="parity", scene=name, res=res, spp=spp, rng=rng, trav=tname,
             tri_equal=bool(np.array_equal(tri, prim["tri"])), tri_mismatch=int((tri != prim["tri"]).sum()),
             k_equal=bool(np.array_equal(k.view(np.uint32), prim["k"].view(np.uint32))),
             pix_identical=float(np.mean(out == ref)), pix_within_1e4=float(np.mean(rel <= 1e-4)),
             max_rel=float(rel.max()), rays_gpu=int(st["rays"]), rays_cpu=int(cnt["rays"]),
             box_gpu=int(st["box_tests"]), box_cpu=int(cnt["box_tests"]), tri_gpu=int(st["tri_tests"]),
             tri_cpu=int(cnt["tri_tests"]), mismatches=int(st["mismatches"]), trace_ms=st["trace_ms"],
             primary_ms=st["primary_ms"], cpu_s=t_cpu, smem=st["scene_in_smem"])


def speed(ctx, name, w, h, spp, bounce, rng=1, trav=0, ibl_name="preview", reps=2):
    sc = fixtures.load_scene(name)
    fixtures.upload(ctx, sc, fixtures.load_ibl(ibl_name))
    cam, env = fixtures.cam_env(sc["params"], w, h)
    o = rt.make_opts(rng_mode=rng, traversal=trav, seed=1)
    for r in range(reps):
        t0 = time.time()
        out = ctx.render(cam, env, w, h, spp, bounce, opts=o)
        wall = time.time() - t0
        st = ctx.stats()
        ms = st["total_ms"]
        emit(kind="speed", scene=name, w=w, h=h, spp=spp, rng=rng, trav=trav, rep=r, total_ms=ms,
             primary_ms=st["primary_ms"], trace_ms=st["trace_ms"], wall_s=wall, rays=int(st["rays"]),
             mrays_s=st["rays"] / ms / 1e3, msamples_s=w * h * spp / ms / 1e3, mean=float(out.mean()),
             smem=st["scene_in_smem"])


def math_checks(ctx):
    rng = np.random.default_rng(0)
    n = 1 << 20
    cases = {
        "sin": (0, rng.uniform(-7, 7, n)), "cos": (1, rng.uniform(-7, 7, n)),
        "acos": (2, rng.uniform(-1, 1, n)), "asin": (3, rng.uniform(-1, 1, n)),
        "tan": (5, rng.uniform(-1.5, 1.5, n)),
    }
    for nm, (fn, a) in cases.items():
        a = a.astype(np.float32)
        g = ctx.math_probe(fn, a)
        c = oracle.math_probe(nm, a)
        emit(kind="math", fn=nm, n=n, mismatches=int((g.view(np.uint32) != c.view(np.uint32)).sum()))
    a = rng.uniform(-1, 1, n).astype(np.float32)
    b = rng.uniform(-1, 1, n).astype(np.float32)
    g = ctx.math_probe(4, a, b)
    c = oracle.math_probe("atan2", a, b)
    emit(kind="math", fn="atan2", n=n, mismatches=int((g.view(np.uint32) != c.view(np.uint32)).sum()))
    a = rng.uniform(0, 1.2, n).astype(np.float32)
    b = np.full(n, 2.2, np.float32)
    g = ctx.math_probe(6, a, b)
    c = oracle.math_probe("pow", a, b)
    emit(kind="math", fn="pow2.2", n=n, mismatches=int((g.view(np.uint32) != c.view(np.uint32)).sum()))
    # the division of the exact slab test vs IEEE division
    a = (rng.standard_normal(n) * 10 ** rng.uniform(-6, 6, n)).astype(np.float32)
    b = (rng.standard_normal(n) * 10 ** rng.uniform(-6, 6, n)).astype(np.float32)
    g = ctx.math_probe(7, a, b)
    with np.errstate(all="ignore"):
        c = (a / b).astype(np.float32)
    emit(kind="math", fn="fdiv", n=n, mismatches=int((g.view(np.uint32) != c.view(np.uint32)).sum()))
    for ctr, k0, k1 in (([0, 0, 0, 0], 0, 0), ([0xffffffff] * 4, 0xffffffff, 0xffffffff),
                        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], 0xa4093822, 0x299f31d0)):
        emit(kind="philox", gpu=[hex(int(v)) for v in ctx.philox_probe(ctr, k0, k1)],
             cpu=[hex(int(v)) for v in oracle.philox(ctr, k0, k1)])


def main():
    ctx = rt.Context(0)
    math_checks(ctx)
    parity(ctx, "cornell", 128, 8, 4)
    parity(ctx, "cornell", 128, 8, 4, rng=1, seed=5)
    parity(ctx, "single", 64, 4, 4)
    parity(ctx, "proto", 96, 4, 4)
    parity(ctx, "furnace", 96, 4, 4)
    parity(ctx, "serre", 96, 4, 4)
    parity(ctx, "monkey_cfg2", 96, 4, 4)
    speed(ctx, "cornell", 512, 512, 16, 4, rng=0)
    speed(ctx, "cornell", 1920, 1080, 64, 4, rng=1)
    speed(ctx, "cornell", 1920, 1080, 64, 4, rng=1, trav=1, reps=1)
    speed(ctx, "monkey_cfg2", 1920, 1080, 32, 4, rng=1)
    speed(ctx, "serre", 1024, 1024, 32, 4, rng=1)
    speed(ctx, "furnace_cfg3", 1024, 1024, 32, 4, rng=1, ibl_name="grey")


if __name__ == "__main__":
    main()
