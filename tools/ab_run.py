#!/usr/bin/env python3
"""Development A/B: time fixture renders with a given build of libb200rt.so.
usage: ab_run.py <lib.so> [scene:w:h:spp ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ensem3a_openclraytracer_b200 import _capi  # noqa: E402

_capi.LIB_PATH = os.path.abspath(sys.argv[1])
if os.environ.get("AB_OLD_ABI"):   # a library built from an older commit: bind only what it exports
    import ctypes
    _probe = ctypes.CDLL(_capi.LIB_PATH)
    _capi.SYMBOLS = [s for s in _capi.SYMBOLS if hasattr(_probe, s)]
import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from tests import fixtures  # noqa: E402


def main():
    cases = sys.argv[2:] or ["cornell:1920:1080:32", "monkey_cfg2:1920:1080:16", "serre:1024:1024:16", "furnace_cfg3:1024:1024:32"]
    ctx = rt.Context(0)
    for c in cases:
        name, w, h, spp = c.split(":")
        w, h, spp = int(w), int(h), int(spp)
        sc = fixtures.load_scene(name)
        fixtures.upload(ctx, sc, fixtures.load_ibl("grey" if name == "furnace_cfg3" else "preview"))
        cam, env = fixtures.cam_env(sc["params"], w, h)
        best = None
        for _ in range(3):
            out = ctx.render(cam, env, w, h, spp, 4, opts=rt.make_opts(rng_mode=1, traversal=0, seed=0))
            st = ctx.stats()
            if best is None or st["total_ms"] < best["total_ms"]:
                best = st
        split = ""
        try:
            ctx.render(cam, env, w, h, min(spp, 2), 4, opts=rt.make_opts(rng_mode=1, traversal=0, seed=0, collect_stats=True))
            st = ctx.stats()
            split = f" | box/ray {st['box_tests'] / st['rays']:.2f} tri/ray {st['tri_tests'] / st['rays']:.2f}"
        except Exception as e:  # noqa: BLE001
            split = f" | stats failed: {e}"
        try:
            ctx.render(cam, env, w, h, spp, 4, opts=rt.make_opts(rng_mode=1, traversal=0, seed=0, time_kernels=True))
            st = ctx.stats()
            split += f" | k_trace {st['trace_kernel_ms']:.2f} k_shade {st['shade_kernel_ms']:.2f} ms"
        except TypeError:
            pass
        print(f"{os.path.basename(sys.argv[1])} {name} {w}x{h} spp{spp}: rays {best['rays']} total {best['total_ms']:.2f} ms "
              f"{best['rays'] / best['total_ms'] / 1e3:.1f} Mrays/s mean {float(out.mean()):.6f} primary {best['primary_ms']:.2f} ms "
              f"launches {best['kernel_launches']} reval {best.get('revalidated')} exact {best.get('exact_walks')}{split}", flush=True)


if __name__ == "__main__":
    main()
