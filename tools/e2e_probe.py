#!/usr/bin/env python3
"""Development probe: where does a multi-GPU end-to-end step spend its wall time?  (torchrun, N ranks)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import ensem3a_openclraytracer_b200 as rt  # noqa: E402
from ensem3a_openclraytracer_b200.multigpu import DistributedRenderer  # noqa: E402
from tests import fixtures  # noqa: E402
import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc, ibl, cam, env = bench.load_workload()
    W, H, spp, mb = 1920, 1080, int(sys.argv[1]) if len(sys.argv) > 1 else 256, 4
    ctx = rt.Context(local)
    fixtures.upload(ctx, sc, ibl)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    dr = DistributedRenderer(ctx, rank, world, reduce=sys.argv[2] if len(sys.argv) > 2 else "peer")
    for it in range(4):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t = [time.perf_counter()]
        ctx.invalidate(); fixtures.upload(ctx, sc, ibl); t.append(time.perf_counter())
        ctx.set_stream(stream.cuda_stream); t.append(time.perf_counter())
        img = dr.render(cam, env, W, H, spp, mb, rng_mode=rt.RNG_PHILOX, seed=0); t.append(time.perf_counter())
        torch.cuda.synchronize(); t.append(time.perf_counter())
        names = ["upload", "set_stream", "render enqueue", "sync"]
        print(f"rank {rank} it {it}: " + ", ".join(f"{n} {1e3 * (b - a):.1f} ms" for n, a, b in zip(names, t, t[1:])), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
