#!/usr/bin/env python3
"""Experiment: how much does ray ordering matter?  Traces the same random rays through k_trace_rays in random order
and sorted by (direction octant, Morton code of the origin); prints both times.
usage: sort_probe.py [scene=monkey_cfg2] [n=4000000]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from tests import fixtures  # noqa: E402


def morton(q, bits):
    out = np.zeros(q.shape[0], np.int64)
    for b in range(bits):
        for k in range(3):
            out |= ((q[:, k] >> b) & 1) << (3 * b + k)
    return out


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "monkey_cfg2"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000000
    sc = fixtures.load_scene(name)
    ctx = rt.Context(0)
    fixtures.upload(ctx, sc)
    bvh = sc["BVH"].reshape(-1, 9)
    lo, hi = bvh[0, 2:5], bvh[0, 5:8]
    r = np.random.default_rng(0)
    o = r.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = r.standard_normal((n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d], 1)
    for bits in (0, 2, 3, 4, 5):
        if bits == 0:
            order = np.arange(n)
            label = "random order"
        else:
            q = np.clip(((o - lo) / (hi - lo) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
            octant = (d[:, 0] > 0).astype(np.int64) | ((d[:, 1] > 0).astype(np.int64) << 1) | ((d[:, 2] > 0).astype(np.int64) << 2)
            key = (octant << (3 * bits)) | morton(q, bits)
            order = np.argsort(key, kind="stable")
            label = f"sorted by octant + {bits}-bit Morton origin"
        best = 1e9
        for _ in range(3):
            ctx.trace_rays(rays[order])
            best = min(best, ctx.stats()["trace_ms"])
        print(f"{name} {n} rays, {label}: {best:.2f} ms  {n / best / 1e3:.0f} Mrays/s", flush=True)


if __name__ == "__main__":
    main()
