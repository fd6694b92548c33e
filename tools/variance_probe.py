#!/usr/bin/env python3
"""Variance at equal spp of the opt-in estimators (b200rt_opts.sampling: 1 glossy importance sampling, 2 light sampling,
3 both) against the reference's, on fixture scenes: RMSE of an spp-sample image against a converged image of the SAME
integrand (the mean of all estimators' long runs), per estimator, with the rays and device time each one spends.  One
JSON line per scene.
usage: variance_probe.py [spp=64] [converged_spp=8192] [scene:w:h:ibl ...]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from tests import fixtures  # noqa: E402


def main():
    a = sys.argv[1:]
    spp = int(a[0]) if a else 64
    big = int(a[1]) if len(a) > 1 else 8192
    cases = a[2:] or ["monkey_cfg2:960:540:preview", "furnace_cfg3:512:512:grey", "serre:960:540:preview"]
    ctx = rt.Context(0)
    for c in cases:
        name, w, h, ibl_name = c.split(":")
        w, h = int(w), int(h)
        sc, ibl = fixtures.load_scene(name), fixtures.load_ibl(ibl_name)
        fixtures.upload(ctx, sc, ibl)
        cam, env = fixtures.cam_env(sc["params"], w, h)

        def sums(mode, seed, n):
            o = rt.make_opts(rng_mode=rt.RNG_PHILOX, seed=seed, sampling=mode, output=rt.OUT_SUMS, sample_streams=-1)
            out = ctx.render(cam, env, w, h, n, 4, opts=o)
            return out.astype(np.float64) / n, ctx.stats()

        modes = ((0, "reference"), (1, "importance"), (2, "lights"), (3, "importance_lights"))
        conv = {m: sums(m, 1000 + m, big)[0] for m, _ in modes}
        truth = sum(conv.values()) / len(conv)
        line = dict(scene=name, width=w, height=h, spp=spp, converged_spp=big,
                    converged_means={k: float(conv[m].mean()) for m, k in modes})
        for mode, key in modes:
            errs, ms, rays = [], [], []
            for seed in range(4):
                img, st = sums(mode, seed, spp)
                errs.append(np.sqrt(np.mean((img - truth) ** 2)))
                ms.append(st["total_ms"])
                rays.append(st["rays"])
            line[f"rmse_{key}"] = float(np.mean(errs))
            line[f"device_ms_{key}"] = float(np.mean(ms))
            line[f"rays_{key}"] = float(np.mean(rays))
        for _, key in modes[1:]:
            r = line["rmse_reference"] / line[f"rmse_{key}"]
            line[f"samples_saved_at_equal_error_{key}"] = r * r
            line[f"time_saved_at_equal_error_{key}"] = r * r * line["device_ms_reference"] / line[f"device_ms_{key}"]
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
