#!/usr/bin/env python3
"""One render of a fixture scene — the command ncu wraps (see profiles/README.md).
usage: profile_run.py <scene> <width> <height> <spp> [rng=1] [traversal=0] [reps=1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from tests import fixtures  # noqa: E402


def main():
    name, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rng = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    trav = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 1
    ctx = rt.Context(0)
    sc = fixtures.load_scene(name)
    fixtures.upload(ctx, sc, fixtures.load_ibl("grey" if name == "furnace_cfg3" else "preview"))
    cam, env = fixtures.cam_env(sc["params"], w, h)
    for _ in range(reps):
        ctx.render(cam, env, w, h, spp, 4, opts=rt.make_opts(rng_mode=rng, traversal=trav, seed=0))
        st = ctx.stats()
        print(f"{name} {w}x{h} spp{spp}: {st['rays']} rays, primary {st['primary_ms']:.3f} ms, paths {st['trace_ms']:.3f} ms, "
              f"{st['rays'] / st['total_ms'] / 1e3:.1f} Mrays/s, smem={st['scene_in_smem']}")


if __name__ == "__main__":
    main()
