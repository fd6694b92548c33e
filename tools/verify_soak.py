#!/usr/bin/env python3
"""Soak run of B200RT_TRAVERSAL_VERIFY: every ray of large renders is traced by the reference-order walk AND by
the fast path on the GPU; prints rays, disagreements (must be 0) per case, and how many fast-path winners needed
the exact re-trace."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from tests import fixtures  # noqa: E402

CASES = [("cornell", 3840, 2160, 32), ("monkey_cfg2", 1920, 1080, 64), ("serre", 3840, 2160, 16),
         ("furnace_cfg3", 1920, 1080, 128), ("proto", 1920, 1080, 64), ("single", 1920, 1080, 16)]


def main():
    ctx = rt.Context(0)
    for name, w, h, spp in CASES:
        sc = fixtures.load_scene(name)
        fixtures.upload(ctx, sc, fixtures.load_ibl("grey" if name == "furnace_cfg3" else "preview"))
        cam, env = fixtures.cam_env(sc["params"], w, h)
        for rng in (rt.RNG_REFERENCE, rt.RNG_PHILOX):
            ctx.render(cam, env, w, h, spp, 4, opts=rt.make_opts(rng_mode=rng, seed=11, traversal=rt.TRAVERSAL_VERIFY, stack_cap=64))
            st = ctx.stats()
            ctx.render(cam, env, w, h, spp, 4, opts=rt.make_opts(rng_mode=rng, seed=11))
            s2 = ctx.stats()
            print(json.dumps(dict(scene=name, width=w, height=h, spp=spp, rng=rng, rays=st["rays"], mismatches=st["mismatches"],
                                  fast_rays=s2["rays"], revalidated=s2["revalidated"])), flush=True)


if __name__ == "__main__":
    main()
