#!/usr/bin/env python3
"""north_star's third tolerance — "the converged image RMSE must be within 1e-3" — for the mode that is actually
timed: the counter-based Philox generator against the reference's own generator (MathLib.cl:294-310, seed = pixel
index), both on the GPU (the reference-generator image is bit-identical to the oracle's and hence to the reference's,
tests/test_gpu_parity.py).  Two independent Philox images give the pure Monte-Carlo noise floor at the same spp; what
the reference-generator image adds on top of it is the bias of its short cycles (SURVEY.md §8a `rand`: every start
value falls into one of a few cycles of 1 814 ... 100 500 draws).
usage: converged_rmse.py [scene=cornell] [res=512] [spp ...]        one JSON line per spp"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from tests import fixtures  # noqa: E402


def rmse(a, b):
    return float(np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))


def main():
    a = sys.argv[1:]
    scene = a[0] if a else "cornell"
    res = int(a[1]) if len(a) > 1 else 512
    spps = [int(x) for x in a[2:]] or [256, 1024, 4096, 16384]
    sc = fixtures.load_scene(scene)
    ctx = rt.Context(0)
    fixtures.upload(ctx, sc, fixtures.load_ibl("grey" if scene == "furnace_cfg3" else "preview"))
    cam, env = fixtures.cam_env(sc["params"], res)
    top = None
    for spp in spps:
        ref = ctx.render(cam, env, res, res, spp, 4, opts=rt.make_opts(rng_mode=rt.RNG_REFERENCE))
        ms_ref = ctx.stats()["total_ms"]
        p0 = ctx.render(cam, env, res, res, spp, 4, opts=rt.make_opts(rng_mode=rt.RNG_PHILOX, seed=0))
        ms_p = ctx.stats()["total_ms"]
        p1 = ctx.render(cam, env, res, res, spp, 4, opts=rt.make_opts(rng_mode=rt.RNG_PHILOX, seed=12345))
        top = (p0.astype(np.float64) + p1.astype(np.float64)) / 2
        line = dict(scene=scene, res=res, spp=spp, rmse_philox_vs_reference_rng=rmse(p0, ref),
                    rmse_philox_vs_philox_other_seed=rmse(p0, p1), rmse_reference_rng_vs_mean_of_two_philox=rmse(ref, top),
                    mean_reference_rng=float(ref.mean()), mean_philox=float(p0.mean()),
                    device_ms_reference_rng=ms_ref, device_ms_philox=ms_p,
                    clamped_fraction=float(np.mean(ref >= 1.0)))
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
