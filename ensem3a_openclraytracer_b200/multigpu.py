"""One frame on N GPUs of one NVSwitch box: one process per GPU (torch.distributed).

Partition: the scene and environment are replicated; every rank traces the whole
frame for its own contiguous SAMPLE RANGE [s0, s1) with the counter-based Philox generator, whose
key (pixel, sample, bounce) makes the image independent of the partition up to float summation
order.  The per-rank float32 partial sums (width*height*3) are the only data that crosses NVLink:

    reduce="nccl"  torch.distributed.reduce(SUM) to rank 0, then the finalize kernel on rank 0;
    reduce="peer"  rank 0 maps every peer's partial-sum buffer (CUDA IPC) and ONE kernel sums them
                   over NVLink P2P loads, divides by spp and clamps (b200rt_reduce_finalize_device);
                   two 4-byte NCCL all-reduces on the same stream act as the device-side barriers
                   (peers finished writing / rank 0 finished reading).

The reference's own generator is one serial stream per pixel (MathLib.cl:294-310) and cannot split
a pixel's samples; with rng_mode=REFERENCE the frame is split by PIXELS instead: rank r takes the rows of
8x4-pixel tiles whose index is r modulo the number of ranks (interleaved, because the cost of a pixel varies by
an order of magnitude between sky and geometry).  Disjoint pixels, so the reduce degenerates to a gather and is
bit-exact.
"""
from . import _capi


def split_range(n, parts):
    """parts contiguous ranges covering [0, n), sizes differing by at most one."""
    base, extra = divmod(int(n), int(parts))
    out, start = [], 0
    for r in range(parts):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def rank_work(rank, world, width, height, spp, rng_mode):
    """(sample_begin, sample_end, tile_row_mod, tile_row_rem) of `rank`; an empty sample range has begin == end,
    tile_row_mod == 0 means every tile row."""
    if rng_mode == _capi.RNG_PHILOX:
        s0, s1 = split_range(spp, world)[rank]
        return s0, s1, 0, 0
    return 0, spp, (world if world > 1 else 0), (rank if world > 1 else 0)


class _RawCudaBuffer:
    """Exposes a b200rt_alloc'ed device buffer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class DistributedRenderer:
    """Renders one frame with all ranks of `group`; the final image lands on rank 0's GPU.

    partial_fn / finalize_fn exist so the partition + reduce logic can be exercised on CPU with the
    gloo backend (tests/test_multigpu_gloo.py injects the oracle); by default they are the CUDA
    kernels and nothing else.
    """

    def __init__(self, ctx, rank, world, reduce="nccl", group=None, partial_fn=None, finalize_fn=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.ctx, self.rank, self.world, self.group = ctx, int(rank), int(world), group
        self.reduce = reduce
        self.partial_fn, self.finalize_fn = partial_fn, finalize_fn
        self.device = device if device is not None else (torch.device("cuda", ctx.device) if ctx is not None else "cpu")
        self._n = 0
        self._accum = self._out = None
        self._raw_ptr = None
        self._peer_ptrs = None
        self._flag = None
        if reduce not in ("nccl", "peer"):
            raise ValueError("reduce must be 'nccl' or 'peer'")
        if reduce == "peer" and partial_fn is not None:
            raise ValueError("reduce='peer' needs the CUDA path")

    # ---- buffers ----------------------------------------------------------------------------------------
    def _ensure(self, n_floats):
        torch = self.torch
        if self._n == n_floats:
            return
        self.release()
        if self.reduce == "peer":
            self._raw_ptr = self.ctx.alloc(n_floats * 4)
            self._accum = torch.as_tensor(_RawCudaBuffer(self._raw_ptr, n_floats), device=self.device)
            handles = [None] * self.world
            self.dist.all_gather_object(handles, self.ctx.ipc_export(self._raw_ptr), group=self.group)
            if self.rank == 0:
                self._peer_ptrs = [self._raw_ptr] + [self.ctx.ipc_open(h) for h in handles[1:]]
            self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        else:
            self._accum = torch.zeros(n_floats, dtype=torch.float32, device=self.device)
        self._out = torch.zeros(n_floats, dtype=torch.float32, device=self.device) if self.rank == 0 else None
        self._n = n_floats

    def release(self):
        if self._peer_ptrs:
            for p in self._peer_ptrs[1:]:
                self.ctx.ipc_close(p)
        self._peer_ptrs = None
        if self._raw_ptr is not None:
            self.torch.cuda.synchronize()
            self.dist.barrier(group=self.group)  # nobody frees while rank 0 may still map it
            self._accum = None
            self.ctx.free(self._raw_ptr)
            self._raw_ptr = None
        self._accum = self._out = None
        self._n = 0

    # ---- one frame -----------------------------------------------------------------------------------------
    def render(self, cam, env, width, height, spp, max_bounce, rng_mode=_capi.RNG_PHILOX, seed=0,
               traversal=_capi.TRAVERSAL_FAST, sample_streams=0):
        """Enqueues the frame; returns rank 0's device tensor (width*height*3, final image) or None."""
        torch, dist = self.torch, self.dist
        n = width * height * 3
        self._ensure(n)
        s0, s1, tmod, trem = rank_work(self.rank, self.world, width, height, spp, rng_mode)
        empty = (s0 == s1) or (tmod > 1 and trem >= (height + 3) // 4)
        if self.partial_fn is not None:
            part = self.partial_fn(cam, env, width, height, spp, max_bounce, rng_mode, seed, s0, s1, tmod, trem)
            self._accum.copy_(torch.as_tensor(part, dtype=torch.float32))
        else:
            self.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
            self._accum.zero_()
            if not empty:
                opts = _capi.make_opts(rng_mode=rng_mode, traversal=traversal, output=_capi.OUT_SUMS, sample_begin=s0,
                                       sample_end=s1, tile_row_mod=tmod, tile_row_rem=trem, seed=seed,
                                       sample_streams=sample_streams)
                self.ctx.render_device(cam, env, width, height, spp, max_bounce, self._accum.data_ptr(), opts)
        if self.reduce == "peer":
            dist.all_reduce(self._flag, group=self.group)       # every rank's partial sums are complete
            if self.rank == 0:
                self.ctx.reduce_finalize_device(self._peer_ptrs, self._out.data_ptr(), width * height, spp)
            dist.all_reduce(self._flag, group=self.group)       # rank 0 has consumed them
            return self._out
        dist.reduce(self._accum, dst=0, op=dist.ReduceOp.SUM, group=self.group)
        if self.rank != 0:
            return None
        if self.finalize_fn is not None:
            self._out.copy_(torch.as_tensor(self.finalize_fn(self._accum.cpu().numpy(), spp)))
        else:
            self.ctx.finalize_device(self._accum.data_ptr(), self._out.data_ptr(), width * height, spp)
        return self._out
