#!/usr/bin/env python3
"""One fixture scene on N GPUs (torchrun), the frame split across ranks — by tile rows with the reference generator
(bit-exact against one GPU) and by sample ranges with Philox — timed against the same frame on one GPU.
Rank 0 prints one JSON line per mode.  Defaults = BASELINE config 4 (Serre_leger, 3840x2160, 8k map).
usage: torchrun --nproc-per-node N tools/multi_run.py [spp=64] [scene=serre] [width=3840] [height=2160] [ibl=8k]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from ensem3a_openclraytracer_b200.multigpu import DistributedRenderer  # noqa: E402
from tests import fixtures  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    a = sys.argv[1:]
    spp = int(a[0]) if len(a) > 0 else 64
    scene = a[1] if len(a) > 1 else "serre"
    W, H = (int(a[2]), int(a[3])) if len(a) > 3 else (3840, 2160)
    ibl_name = a[4] if len(a) > 4 else "8k"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc = fixtures.load_scene(scene)
    ibl = fixtures.load_ibl(ibl_name)
    cam, env = fixtures.cam_env(sc["params"], W, H)
    ctx = rt.Context(local)
    fixtures.upload(ctx, sc, ibl)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    for mode, name in ((rt.RNG_REFERENCE, "tile rows, reference RNG"), (rt.RNG_PHILOX, "sample ranges, Philox")):
        dr = DistributedRenderer(ctx, rank, world, reduce="peer")
        img = None
        ms = []
        for rep in range(3):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            img = dr.render(cam, env, W, H, spp, 4, rng_mode=mode, seed=0)
            b.record(stream)
            torch.cuda.synchronize(); dist.barrier()
            ms.append(a.elapsed_time(b))
        t = torch.tensor([min(ms)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        line = None
        if rank == 0:
            multi = img.cpu().numpy()
            ctx.set_stream(None)
            single = ctx.render(cam, env, W, H, spp, 4, opts=rt.make_opts(rng_mode=mode, seed=0))
            st = ctx.stats()
            ctx.set_stream(stream.cuda_stream)
            rel = np.abs(multi - single) / np.maximum(np.abs(single), 1e-3)
            line = dict(scene=scene, split=name, n_gpus=world, width=W, height=H, spp=spp, ms=float(t.item()),
                        single_gpu_ms=st["total_ms"], speedup=st["total_ms"] / float(t.item()),
                        mrays_s=st["rays"] / float(t.item()) / 1e3, identical_frac=float(np.mean(multi == single)),
                        max_rel=float(rel.max()), rmse=float(np.sqrt(np.mean((multi - single) ** 2))))
            print(json.dumps(line), flush=True)
        dist.barrier()
        dr.release()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
