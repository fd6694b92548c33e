#!/usr/bin/env python3
"""The reference-facing call on several GPUs WITHOUT torchrun: KernelLauncher(cuda_devices=[0..N-1]).launch_Raytracing —
one host process, the multi-GPU handle of the C ABI (b200rt_multi_*) — timed end to end with host buffers (wall clock
around the call: uploads are cached after the first call, the image is read back every call) and compared with the
same call on one GPU.  One JSON line per (scene, generator).
usage: multi_handle_run.py <n_gpus> [scene:w:h:spp ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from tests import fixtures  # noqa: E402


def main():
    n = int(sys.argv[1])
    cases = sys.argv[2:] or ["monkey_cfg2:1920:1080:256", "cornell:1920:1080:256"]
    one = rt.KernelLauncher(None, None, None, None, cuda_device=0)
    many = rt.KernelLauncher(None, None, None, None, cuda_devices=list(range(n)))
    for c in cases:
        name, w, h, spp = c.split(":")
        w, h, spp = int(w), int(h), int(spp)
        sc = fixtures.load_scene(name)
        ibl = fixtures.load_ibl("8k" if name == "serre" else "preview")
        cam, env = fixtures.cam_env(sc["params"], w, h)
        args = (sc["V_p"], sc["V_n"], sc["V_uv"], sc["faceData"], sc["materialData"], sc["lightData"], sc["BVH"], cam, env,
                w * h, spp, 4, ibl)
        for rng, label in ((rt.RNG_PHILOX, "philox, sample ranges"), (rt.RNG_REFERENCE, "reference generator, tile rows")):
            res = {}
            for kl, key in ((one, "one_gpu"), (many, "n_gpus")):
                kl.rng_mode, kl.seed, kl.sample_streams = rng, 0, -1   # sample streams apply to Philox only
                out = np.zeros(w * h * 3, np.float32)
                kl.launch_Raytracing(out, *args)                       # uploads + warm-up
                best = 1e9
                for _ in range(2):
                    t0 = time.perf_counter()
                    kl.launch_Raytracing(out, *args)
                    best = min(best, time.perf_counter() - t0)
                res[key] = (best, out, kl.last_stats)
            a, b = res["one_gpu"][1], res["n_gpus"][1]
            rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-3)
            rays = res["one_gpu"][2]["rays"]
            print(json.dumps(dict(scene=name, width=w, height=h, spp=spp, split=label, n_gpus=n,
                                  one_gpu_call_s=res["one_gpu"][0], n_gpu_call_s=res["n_gpus"][0],
                                  speedup=res["one_gpu"][0] / res["n_gpus"][0], mrays_s_e2e=rays / res["n_gpus"][0] / 1e6,
                                  device_ms_n_gpus=res["n_gpus"][2]["total_ms"], identical_frac=float(np.mean(a == b)),
                                  max_rel=float(rel.max()), rmse=float(np.sqrt(np.mean((a - b) ** 2))))), flush=True)
    one.close()
    many.close()


if __name__ == "__main__":
    main()
