"""CPU, world_size 2, gloo: the multi-GPU partition + reduce + finalize logic of
ensem3a_openclraytracer_b200.multigpu with the oracle standing in for each rank's GPU
(partial_fn / finalize_fn injection — test-only; the product defaults are the CUDA kernels)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ensem3a_openclraytracer_b200 import multigpu  # noqa: E402
from ensem3a_openclraytracer_b200 import _capi  # noqa: E402


def test_split_range_covers_everything():
    for n in (0, 1, 7, 256, 1000):
        for parts in (1, 2, 3, 8):
            r = multigpu.split_range(n, parts)
            assert len(r) == parts and r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_rank_work_partitions():
    W, H, spp = 50, 30, 10
    s = [multigpu.rank_work(r, 4, W, H, spp, _capi.RNG_PHILOX) for r in range(4)]
    assert [x[:2] for x in s] == [(0, 3), (3, 6), (6, 8), (8, 10)] and all(x[2:] == (0, 0) for x in s)
    p = [multigpu.rank_work(r, 4, W, H, spp, _capi.RNG_REFERENCE) for r in range(4)]
    assert all(x[:2] == (0, spp) for x in p) and [x[2:] for x in p] == [(4, 0), (4, 1), (4, 2), (4, 3)]
    assert multigpu.rank_work(0, 1, W, H, spp, _capi.RNG_REFERENCE) == (0, spp, 0, 0)


def reference_finalize(sums, spp):
    """Raytracing.cl:211-219 in numpy (fmin/fmax drop NaNs)."""
    s = np.asarray(sums, dtype=np.float32) / np.float32(spp)
    return np.fmax(np.fmin(s, np.float32(1.0)), np.float32(0.0)).astype(np.float32)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rng_mode, q):
    from oracle import oracle
    from tests import fixtures
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = fixtures.load_scene("cornell")
    ibl = fixtures.load_ibl()
    W, H, spp, mb = 40, 24, 6, 3
    cam, env = fixtures.cam_env(sc["params"], W, H)

    def partial_fn(cam, env, width, height, spp, max_bounce, rng, seed, s0, s1, tmod, trem):
        out = np.zeros(width * height * 3, np.float32)
        if s0 == s1:
            return out
        tile_rows = range((height + 3) // 4)
        for ty in (tile_rows if tmod <= 1 else [t for t in tile_rows if t % tmod == trem]):
            i0, i1 = ty * 4 * width, min((ty + 1) * 4, height) * width      # one row of 8x4-pixel tiles
            part, _ = oracle.render(sc, cam, env, width * height, spp, max_bounce, ibl, i0=i0, i1=i1, rng_mode=rng,
                                    seed=seed, s0=s0, s1=s1, raw_sums=True, nthreads=2)
            out[3 * i0:3 * i1] = part[3 * i0:3 * i1]
        return out

    dr = multigpu.DistributedRenderer(None, rank, world, reduce="nccl", partial_fn=partial_fn,
                                      finalize_fn=reference_finalize, device="cpu")
    img = dr.render(cam, env, W, H, spp, mb, rng_mode=rng_mode, seed=9)
    if rank == 0:
        full, _ = oracle.render(sc, cam, env, W * H, spp, mb, ibl, rng_mode=rng_mode, seed=9, nthreads=2)
        q.put((img.numpy().copy(), full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("rng_mode", [_capi.RNG_PHILOX, _capi.RNG_REFERENCE])
def test_two_ranks_reproduce_the_single_rank_image(rng_mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rng_mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    img, full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if rng_mode == _capi.RNG_REFERENCE:
        # pixel partition: disjoint pixels, x + 0 is exact -> bit-identical to one rank doing it all
        assert np.array_equal(img.view(np.uint32), full.view(np.uint32))
    else:
        # sample partition: same samples, different float summation order
        np.testing.assert_allclose(img, full, rtol=1e-5, atol=1e-6)
