// b200rt kernels (sm_100a):
//   k_primary   one thread per pixel: camera ray + closest hit; stores the hit (the parity artefact)
//               and finishes every pixel whose samples need no further ray (primary miss or emitter).
//   k_paths     persistent path tracer: each lane owns one pixel at a time and walks its samples as a
//               state machine (shade -> trace -> resolve) so that lanes at different bounces and
//               different samples share the same traversal loop; lanes that run dry are refilled
//               with a warp-aggregated fetch (ballot + one atomic per warp).
//   k_finalize / k_reduce_finalize / k_img_processing   the accumulate / clamp / tonemap passes.
#pragma once
#include "rt_shade.cuh"
#include "rt_trace.cuh"

namespace b200rt {

constexpr int kBlock = 128;

struct DeviceCounters {
  unsigned long long rays, box_tests, tri_tests, mismatches, samples;
};

struct KernelArgs {
  FrameParams F;
  SceneView S;            // global-memory view (the SMEM variant rebuilds node/tri/normal pointers)
  cudaTextureObject_t ibl;
  float4 *prim_dirk;      // per pixel: primary direction xyz, hit distance
  int *prim_tri;          // per pixel: triangle whose samples still need tracing, or -1
  float *out;             // width*height*3
  unsigned int *work_counter;
  DeviceCounters *counters;
  int n_nodes, n_tris;    // repacked counts (for staging)
  int stack_depth;        // entries per lane
  int n_work;             // work items (32-pixel tiles x 32)
  int tiles_x;
  int quorum;             // k_paths leaves its traversal loop when fewer lanes than this are still traversing
  int refill_min;         // idle lanes pull new rays once this many are idle
  int tri_quorum;         // parked triangles are tested once this many lanes hold one
  int slots_per_lane;     // path slots per lane in the warp's shared-memory pool
};

// ---- shared-memory staging of a small scene with the bulk-copy engine (TMA 1-D) ------------------------
RT_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

RT_DEV void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// copies nodes / tris / normals into shared memory; returns a view that points there
template <bool SMEM>
RT_DEV SceneView stage_scene(const KernelArgs &A, unsigned char *smem, size_t *used) {
  SceneView S = A.S;
  if (!SMEM) {
    *used = 0;
    return S;
  }
  __shared__ __align__(8) uint64_t bar;
  const uint32_t nb = (uint32_t)A.n_nodes * 64u, tb = (uint32_t)A.n_tris * 48u, nn = (uint32_t)A.n_tris * 16u;
  float4 *s_nodes = reinterpret_cast<float4 *>(smem);
  float4 *s_tris = reinterpret_cast<float4 *>(smem + nb);
  float4 *s_nrm = reinterpret_cast<float4 *>(smem + nb + tb);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nb + tb + nn)
                 : "memory");
    if (nb) bulk_g2s(s_nodes, A.S.nodes, nb, &bar);
    bulk_g2s(s_tris, A.S.tris, tb, &bar);
    bulk_g2s(s_nrm, A.S.normals, nn, &bar);
  }
  if (threadIdx.x == 0) {  // one thread waits for the bytes to land; the block barrier publishes them
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(&bar))
          : "memory");
    }
  }
  __syncthreads();
  S.nodes = s_nodes;
  S.tris = s_tris;
  S.normals = s_nrm;
  *used = (size_t)nb + tb + nn;
  return S;
}

RT_DEV int work_to_pixel(const KernelArgs &A, unsigned int w) {
  // 32 consecutive work items = one 8x4 pixel tile; tiles in row-major order
  unsigned int tile = w >> 5, in = w & 31u;
  int tx = (int)(tile % (unsigned)A.tiles_x), ty = (int)(tile / (unsigned)A.tiles_x);
  int x = tx * 8 + (int)(in & 7u), y = ty * 4 + (int)(in >> 3);
  if (x >= A.F.width || y >= A.F.height) return -1;
  int i = y * A.F.width + x;
  if (i < A.F.pixel_begin || i >= A.F.pixel_end) return -1;
  return i;
}

RT_DEV void write_pixel(const KernelArgs &A, int i, v3 sum) {
  float *o = A.out + 3 * (size_t)i;
  if (A.F.out_mode == 1) {
    o[0] = sum.x; o[1] = sum.y; o[2] = sum.z;
  } else {  // Raytracing.cl:211-219
    float n = (float)A.F.spp;
    o[0] = clamp01(sum.x / n);
    o[1] = clamp01(sum.y / n);
    o[2] = clamp01(sum.z / n);
  }
}

// ---- primary hits ---------------------------------------------------------------------------------------
// PARITY: write tri / k for every pixel of [pixel_begin,pixel_end) and nothing else.
template <int TRAV, bool SMEM, bool PARITY, bool STATS>
__global__ void __launch_bounds__(kBlock) k_primary(const __grid_constant__ KernelArgs A, int *tri_out, float *k_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  size_t used;
  SceneView S = stage_scene<SMEM>(A, smem, &used);
  LaneStack st;
  st.base = reinterpret_cast<float2 *>(smem + used) + threadIdx.x;
  st.stride = kBlock;
  TraceCounters tc;
  tc.box_tests = 0; tc.tri_tests = 0;
  unsigned int mism = 0;
  unsigned long long rays = 0;
  const unsigned int n_work = (unsigned)A.n_work;
  for (unsigned int w = blockIdx.x * kBlock + threadIdx.x; w < n_work; w += gridDim.x * kBlock) {
    int i = work_to_pixel(A, w);
    if (i < 0) continue;
    v3 d = camera_dir(A.F, i);
    Hit h = closest_hit<TRAV, SMEM, STATS>(S, A.F.cam_pos, d, st, &tc, &mism);
    rays++;
    if (PARITY) {
      tri_out[i] = h.tri;
      k_out[i] = h.k;
      continue;
    }
    A.prim_dirk[i] = make_float4(d.x, d.y, d.z, h.k);
    int mat_type = -1;
    float emit = 0.0f;
    if (h.tri >= 0) {
      int mat = __float_as_int(ld4<SMEM>(S.tris + 3 * (size_t)h.tri + 2).y);
      Material m = load_material(S.mats, mat);
      mat_type = m.type;
      emit = m.roughness;
    }
    if (h.tri >= 0 && mat_type != 0) {
      A.prim_tri[i] = h.tri;  // k_paths takes it from here
      continue;
    }
    // every sample of this pixel is the same ray-free value: Raytracing.cl:146-150 (miss) / :140-144 (emitter)
    v3 c;
    if (h.tri < 0) {
      c = (mk3(1.0f, 1.0f, 1.0f) * ibl_lookup(A.F, A.ibl, d)) * A.F.ibl_power;
    } else {
      c = mk3(1.0f, 1.0f, 1.0f) * emit;
    }
    v3 sum = mk3(0.0f, 0.0f, 0.0f);
    for (int s = A.F.s0; s < A.F.s1; ++s) sum = sum + c;  // the reference accumulates sample by sample
    write_pixel(A, i, sum);
    A.prim_tri[i] = -1;
  }
  // one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    rays += __shfl_down_sync(0xffffffffu, rays, o);
    mism += __shfl_down_sync(0xffffffffu, mism, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (rays) atomicAdd(&A.counters->rays, rays);
    if (mism) atomicAdd(&A.counters->mismatches, (unsigned long long)mism);
  }
  if (STATS) {
    if (tc.box_tests) atomicAdd(&A.counters->box_tests, tc.box_tests);
    if (tc.tri_tests) atomicAdd(&A.counters->tri_tests, tc.tri_tests);
  }
}

// ---- path tracer ----------------------------------------------------------------------------------------------
// One warp = one small wavefront renderer.  Each warp owns NS = 32 * slots_per_lane PATH SLOTS in shared
// memory (structure of arrays, one pixel per slot, samples of a pixel strictly in order; the running sum of a
// pixel lives in the output buffer).  The warp alternates between two phases:
//
//   phase A  (shade)     slots whose trace has finished, or that start a sample, are compacted into a list
//                        with ballot/popc prefix sums and processed 32 at a time: resolve the hit, end/start
//                        the sample, sample the next direction (Raytracing.cl:51-137).  Every processed slot
//                        ends with a ray READY to trace (or becomes FREE when its pixel is complete).
//   phase B  (traverse)  lanes pull READY rays from the compacted ray list and advance their traversals one
//                        node per turn.  Leaves are parked and tested by the whole warp once `tri_quorum`
//                        lanes hold one.  A lane whose ray is done hands the hit back to its slot and — as
//                        soon as `refill_min` lanes are idle — the idle lanes pull the next rays together.
//                        When the list is empty the warp keeps stepping until fewer than `quorum` lanes are
//                        left, parks the unfinished traversals in shared memory and returns to phase A.
//
// FREE slots take new pixels from a global work counter (one warp-aggregated atomic per layer).
enum SlotField { F_PIX = 0, F_META, F_RNG, F_OX, F_OY, F_OZ, F_DX, F_DY, F_DZ, F_K, F_TRI, F_AX, F_AY, F_AZ, F_COUNT };
// meta word: bits 0-2 state, bit 3 sun ray, bits 4-7 material type of the surface the ray left, bits 8-15 bounce j,
// bits 16-31 sample index
enum SlotState { ST_FREE = 0, ST_START = 1, ST_SHADE = 2, ST_READY = 3, ST_FLIGHT = 4, ST_DONE = 5 };
RT_DEV uint32_t meta_pack(int state, int sun, int type, int j, int s) {
  return (uint32_t)state | ((uint32_t)sun << 3) | ((uint32_t)(type & 15) << 4) | ((uint32_t)(j & 255) << 8) | ((uint32_t)s << 16);
}
constexpr int kTravSaveWords = 11 + kParkCap;  // o, d, best.k, best.tri, best_rank, cur, sp  +  parked leaves

__host__ __device__ inline size_t path_pool_bytes_per_warp(int slots_per_lane) {
  return ((size_t)F_COUNT * 4u + 2u) * 32u * (size_t)slots_per_lane;  // fields + two uint8 lists
}
__host__ __device__ inline size_t path_extra_smem_bytes(int slots_per_lane) {  // beyond scene + stacks
  return (size_t)kBlock * kTravSaveWords * 4u + (size_t)(kBlock / 32) * path_pool_bytes_per_warp(slots_per_lane);
}

template <int TRAV, bool SMEM, bool STATS>
__global__ void __launch_bounds__(kBlock) k_paths(const __grid_constant__ KernelArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  size_t used;
  const SceneView S = stage_scene<SMEM>(A, smem, &used);
  LaneStack st;
  st.base = reinterpret_cast<float2 *>(smem + used) + threadIdx.x;
  st.stride = kBlock;
  const FrameParams &F = A.F;
  const unsigned int lane = threadIdx.x & 31u;
  const unsigned int lt_mask = (1u << lane) - 1u;
  const int K = A.slots_per_lane;
  const int NS = 32 * K;
  uint32_t *tsave = reinterpret_cast<uint32_t *>(smem + used + (size_t)kBlock * A.stack_depth * sizeof(float2)) + threadIdx.x;
  uint32_t *pool = reinterpret_cast<uint32_t *>(smem + used + (size_t)kBlock * A.stack_depth * sizeof(float2) +
                                                (size_t)kBlock * kTravSaveWords * 4u) +
                   (threadIdx.x >> 5) * (path_pool_bytes_per_warp(K) / 4);
  unsigned char *alist = reinterpret_cast<unsigned char *>(pool + F_COUNT * NS);
  unsigned char *rlist = alist + NS;
#define FLD(f, slot) pool[(f) * NS + (slot)]
#define FLDF(f, slot) __uint_as_float(pool[(f) * NS + (slot)])

  for (int k = 0; k < K; ++k) {
    FLD(F_PIX, k * 32 + lane) = 0xffffffffu;
    FLD(F_META, k * 32 + lane) = meta_pack(ST_FREE, 0, 0, 0, 0);
  }
  __syncwarp();

  int my_slot = -1;      // slot whose ray this lane is tracing (its traversal is parked in tsave during phase A)
  bool exhausted = false;
  int pendingA = 0;      // slots waiting for phase A (uniform)

  unsigned long long rays = 0, samples = 0;
  unsigned int mism = 0;
  TraceCounters tc;
  tc.box_tests = 0; tc.tri_tests = 0;

  for (;;) {
    // ===================================== phase A ==========================================================
    // A0: FREE slots fetch pixels; a pixel that k_primary already finished leaves the slot FREE for the next round
    for (;;) {
      bool any_free = false;
      for (int k = 0; k < K && !exhausted; ++k) {
        const int slot = k * 32 + (int)lane;
        const bool is_free = (FLD(F_META, slot) & 7u) == ST_FREE;
        const unsigned int need = __ballot_sync(0xffffffffu, is_free);
        if (need == 0u) continue;
        unsigned int base = 0;
        const int leader = __ffs(need) - 1;
        const unsigned int cnt = __popc(need);
        if ((int)lane == leader) base = atomicAdd(A.work_counter, cnt);
        base = __shfl_sync(0xffffffffu, base, leader);
        if (base + cnt >= (unsigned)A.n_work) exhausted = true;
        bool still_free = is_free;
        if (is_free) {
          unsigned int w = base + __popc(need & lt_mask);
          if (w < (unsigned)A.n_work) {
            int i = work_to_pixel(A, w);
            if (i >= 0 && A.prim_tri[i] >= 0) {
              FLD(F_PIX, slot) = (uint32_t)i;
              FLD(F_RNG, slot) = (uint32_t)i;  // Raytracing.cl:171 (imgSize receives imgDim, so seed0 = i)
              FLD(F_META, slot) = meta_pack(ST_START, 0, 0, 0, F.s0);
              float *acc_px = A.out + 3 * (size_t)i;  // the pixel's running sum lives in the output buffer
              acc_px[0] = 0.0f; acc_px[1] = 0.0f; acc_px[2] = 0.0f;
              still_free = false;
            }
          }
        }
        if (__ballot_sync(0xffffffffu, still_free) != 0u) any_free = true;
      }
      if (!any_free || exhausted) break;
    }
    __syncwarp();

    // A1: compact the slots that need shading work into alist
    int n_a = 0;
    for (int k = 0; k < K; ++k) {
      const int slot = k * 32 + (int)lane;
      const unsigned int stt = FLD(F_META, slot) & 7u;
      const bool want = (stt == ST_START) || (stt == ST_SHADE) || (stt == ST_DONE);
      const unsigned int m = __ballot_sync(0xffffffffu, want);
      if (want) alist[n_a + __popc(m & lt_mask)] = (unsigned char)slot;
      n_a += __popc(m);
    }
    __syncwarp();

    // A2: process them 32 at a time; every processed slot ends READY (ray to trace) or FREE (pixel complete)
    int n_ready = 0;
    for (int c0 = 0; c0 < n_a; c0 += 32) {
      const bool have = c0 + (int)lane < n_a;
      bool ready = false;
      int slot = 0;
      if (have) {
        slot = alist[c0 + lane];
        const uint32_t meta = FLD(F_META, slot);
        int state = (int)(meta & 7u), sun_ray = (int)((meta >> 3) & 1u), seg_type = (int)((meta >> 4) & 15u);
        int j = (int)((meta >> 8) & 255u), s = (int)(meta >> 16);
        const int pix = (int)FLD(F_PIX, slot);
        v3 o = mk3(FLDF(F_OX, slot), FLDF(F_OY, slot), FLDF(F_OZ, slot));   // ray origin (DONE) / segment origin
        v3 d = mk3(FLDF(F_DX, slot), FLDF(F_DY, slot), FLDF(F_DZ, slot));   // bounce direction; for a sun ray: the escaped direction
        float hk = FLDF(F_K, slot);
        int htri = (int)FLD(F_TRI, slot);
        v3 acc = mk3(FLDF(F_AX, slot), FLDF(F_AY, slot), FLDF(F_AZ, slot));
        bool pixel_done = false;

        if (state == ST_DONE) {  // ---- resolve the finished trace
          bool end_sample = false;
          if (!sun_ray) {
            if (htri >= 0) {  // the bounce ray becomes the current segment (:91-93)
              int mat = __float_as_int(ld4<SMEM>(S.tris + 3 * (size_t)htri + 2).y);
              Material mb = load_material(S.mats, mat);
              if (mb.type != 0) {
                if (j == F.max_bounce) {  // :99-103
                  acc = mk3(0.0f, 0.0f, 0.0f);
                  end_sample = true;
                } else {
                  ++j;
                  state = ST_SHADE;
                }
              } else {  // :105-109
                acc = acc * mb.roughness;
                end_sample = true;
              }
            } else {  // escaped: shadow ray towards the sun from the same origin (:115-124); d keeps the escaped direction
              sun_ray = 1;
              state = ST_READY;
            }
          } else {  // sun ray finished, :125-137
            v3 sun = mk3(0.0f, 0.0f, 0.0f);
            if (htri < 0) {
              if (seg_type != 3) sun = mk3(F.sun_power, F.sun_power, F.sun_power);
            } else {
              int mat = __float_as_int(ld4<SMEM>(S.tris + 3 * (size_t)htri + 2).y);
              Material ms = load_material(S.mats, mat);
              if (ms.type == 3) sun = ms.color * F.sun_power;
            }
            v3 envl = ibl_lookup(F, A.ibl, d) * F.ibl_power;
            acc = acc * (sun + envl);
            end_sample = true;
          }
          if (end_sample) {
            float *acc_px = A.out + 3 * (size_t)pix;
            v3 sum = mk3(acc_px[0], acc_px[1], acc_px[2]);
            sum = sum + acc;  // :207
            ++s;
            ++samples;
            if (s >= F.s1) {
              write_pixel(A, pix, sum);
              pixel_done = true;
            } else {
              acc_px[0] = sum.x; acc_px[1] = sum.y; acc_px[2] = sum.z;
              state = ST_START;
            }
          }
        }

        if (!pixel_done && state == ST_START) {  // ---- reload the cached primary hit (Raytracing.cl:195-201)
          float4 dk = A.prim_dirk[pix];
          o = F.cam_pos;
          d = mk3(dk.x, dk.y, dk.z);
          hk = dk.w;
          htri = A.prim_tri[pix];
          acc = mk3(1.0f, 1.0f, 1.0f);
          j = 0;
          state = ST_SHADE;
        }

        if (!pixel_done && state == ST_SHADE) {  // ---- next direction, BRDF * cos / pdf (:51-87); segment = (o, d, hk, htri)
          float4 t2 = ld4<SMEM>(S.tris + 3 * (size_t)htri + 2);
          float4 nn = ld4<SMEM>(S.normals + htri);
          v3 n = mk3(nn.x, nn.y, nn.z);
          Material m = load_material(S.mats, __float_as_int(t2.y));
          v3 nd, brdf;
          float inv_pdf;
          if (m.type == 3) {
            nd = d;
            brdf = m.color;
            inv_pdf = 1.0f / fabsf(dot(nd, unit(n)));
          } else {
            float u0, u1;
            rng_state g;
            g.a = FLD(F_RNG, slot);
            if (F.rng_mode == 0) draw2<0>(g, (uint32_t)pix, (uint32_t)s, (uint32_t)j, F.key0, F.key1, &u0, &u1);
            else draw2<1>(g, (uint32_t)pix, (uint32_t)s, (uint32_t)j, F.key0, F.key1, &u0, &u1);
            FLD(F_RNG, slot) = g.a;
            if (m.type == 1) {
              nd = sample_cosine(n, u0, u1, &inv_pdf);
              brdf = m.color * (1.0f / 3.14f);
            } else {
              nd = sample_uniform(n, u0, u1, &inv_pdf);
              brdf = bsdf_ggx(m, neg3(d), nd, n);
            }
          }
          o = o + unit(d) * hk;  // :79 — no offset along the normal
          d = nd;
          float att = inv_pdf * fabsf(dot(nd, unit(n)));
          acc = (acc * brdf) * att;
          seg_type = m.type;
          sun_ray = 0;
          state = ST_READY;
        }

        if (pixel_done) {
          FLD(F_PIX, slot) = 0xffffffffu;
          FLD(F_META, slot) = meta_pack(ST_FREE, 0, 0, 0, 0);
        } else {
          FLD(F_OX, slot) = __float_as_uint(o.x); FLD(F_OY, slot) = __float_as_uint(o.y); FLD(F_OZ, slot) = __float_as_uint(o.z);
          FLD(F_DX, slot) = __float_as_uint(d.x); FLD(F_DY, slot) = __float_as_uint(d.y); FLD(F_DZ, slot) = __float_as_uint(d.z);
          FLD(F_AX, slot) = __float_as_uint(acc.x); FLD(F_AY, slot) = __float_as_uint(acc.y); FLD(F_AZ, slot) = __float_as_uint(acc.z);
          FLD(F_META, slot) = meta_pack(state, sun_ray, seg_type, j, s);
          ready = (state == ST_READY);
        }
      }
      const unsigned int rm = __ballot_sync(0xffffffffu, ready);
      if (ready) rlist[n_ready + __popc(rm & lt_mask)] = (unsigned char)slot;
      n_ready += __popc(rm);
    }
    pendingA = 0;
    __syncwarp();

    // nothing to trace, nothing in flight: either everything is finished or only FREE slots remain
    if (n_ready == 0 && __ballot_sync(0xffffffffu, my_slot >= 0) == 0u) {
      if (exhausted) break;
      continue;
    }

    // ===================================== phase B ==========================================================
    Trav T;
    int pn = 0;                                  // parked leaves of this lane
    uint32_t *parks = tsave + 11 * kBlock;       // entry e at parks[e * kBlock]
    T.active = false;
    T.best.tri = -1; T.best.k = 1000.0f; T.best_rank = 0x7fffffff; T.cur = 0; T.sp = 0;
    T.R.o = mk3(0, 0, 0); T.R.d = mk3(1, 1, 1); T.R.r = mk3(1, 1, 1); T.R.fast = false;
    if (my_slot >= 0) {  // un-park the traversal this lane left unfinished
      v3 o = mk3(__uint_as_float(tsave[0 * kBlock]), __uint_as_float(tsave[1 * kBlock]), __uint_as_float(tsave[2 * kBlock]));
      v3 d = mk3(__uint_as_float(tsave[3 * kBlock]), __uint_as_float(tsave[4 * kBlock]), __uint_as_float(tsave[5 * kBlock]));
      T.R = make_raydiv(o, d, S.fast_div_ok != 0);
      T.best.k = __uint_as_float(tsave[6 * kBlock]);
      T.best.tri = (int)tsave[7 * kBlock];
      T.best_rank = (int)tsave[8 * kBlock];
      T.cur = (int)tsave[9 * kBlock];
      T.sp = (int)tsave[10 * kBlock];
      T.active = true;
    }
    int r_head = 0;
    unsigned int last_nodem = 0xffffffffu;  // forces the bookkeeping path on the first turn
    bool leave = false;
    while (!leave) {
      const bool can_node = T.active && pn <= kParkCap - 2;
      const unsigned int nodem = __ballot_sync(0xffffffffu, can_node);
      const unsigned int parkm = __ballot_sync(0xffffffffu, pn > 0);
      const int n_park = __popc(parkm);
      if (nodem == last_nodem && nodem != 0u && n_park < A.tri_quorum) {  // nothing changed: just step
        if (can_node) trav_step_park<SMEM, STATS>(S, T, pn, parks, kBlock, st, &tc);
        continue;
      }
      // ---- parked triangles: one round, every lane that holds one tests its most recent ------------------------
      const int n_node = __popc(nodem);
      if (n_park >= A.tri_quorum || (n_park > 0 && n_node < A.quorum)) {
        if (pn > 0) {
          --pn;
          if (STATS) tc.tri_tests++;
          test_triangle<SMEM>(S, (int)parks[pn * kBlock], T.R.o, T.R.d, T.best, T.best_rank);
        }
        last_nodem = 0xffffffffu;
        continue;
      }
      // ---- hand finished hits back to their slots (a finished lane holds no parked triangle here) -----------------
      const bool finished = (my_slot >= 0) && !T.active && pn == 0;
      if (finished) {
        FLD(F_K, my_slot) = __float_as_uint(T.best.k);
        FLD(F_TRI, my_slot) = (uint32_t)T.best.tri;
        FLD(F_META, my_slot) = (FLD(F_META, my_slot) & ~7u) | ST_DONE;
        my_slot = -1;
      }
      pendingA += __popc(__ballot_sync(0xffffffffu, finished));
      // ---- idle lanes pull the next rays together ------------------------------------------------------------------
      const unsigned int idle = __ballot_sync(0xffffffffu, my_slot < 0);
      const int n_idle = __popc(idle);
      const int avail = n_ready - r_head;
      if (avail > 0 && (n_idle >= A.refill_min || n_idle == 32 || n_node < A.quorum)) {
        const int rank = __popc(idle & lt_mask);
        if (my_slot < 0 && rank < avail) {
          my_slot = rlist[r_head + rank];
          const uint32_t meta = FLD(F_META, my_slot);
          FLD(F_META, my_slot) = (meta & ~7u) | ST_FLIGHT;
          v3 o = mk3(FLDF(F_OX, my_slot), FLDF(F_OY, my_slot), FLDF(F_OZ, my_slot));
          v3 d = ((meta >> 3) & 1u) ? F.sun_dir : mk3(FLDF(F_DX, my_slot), FLDF(F_DY, my_slot), FLDF(F_DZ, my_slot));
          rays++;
          if (TRAV == 0) {
            trav_begin<SMEM, STATS>(S, T, o, d, &tc);
          } else {  // reference / verify traversal: the whole walk at once
            T.best = closest_hit<TRAV, SMEM, STATS>(S, o, d, st, &tc, &mism);
            T.active = false;
          }
        }
        r_head += min(n_idle, avail);
        last_nodem = 0xffffffffu;  // lanes whose ray ended at the root box are handed back on the next turn
        continue;
      }
      // ---- leave or keep stepping ---------------------------------------------------------------------------------------
      if (n_node == 0) {
        // no lane can step and nothing is parked (else the flush above ran): all remaining lanes are idle
        leave = true;
      } else if (avail == 0 && n_node < A.quorum && pendingA > 0) {
        leave = true;
      } else {
        last_nodem = nodem;
      }
    }
    // park the unfinished traversals for the duration of phase A
    if (my_slot >= 0) {
      tsave[0 * kBlock] = __float_as_uint(T.R.o.x); tsave[1 * kBlock] = __float_as_uint(T.R.o.y); tsave[2 * kBlock] = __float_as_uint(T.R.o.z);
      tsave[3 * kBlock] = __float_as_uint(T.R.d.x); tsave[4 * kBlock] = __float_as_uint(T.R.d.y); tsave[5 * kBlock] = __float_as_uint(T.R.d.z);
      tsave[6 * kBlock] = __float_as_uint(T.best.k);
      tsave[7 * kBlock] = (uint32_t)T.best.tri;
      tsave[8 * kBlock] = (uint32_t)T.best_rank;
      tsave[9 * kBlock] = (uint32_t)T.cur;
      tsave[10 * kBlock] = (uint32_t)T.sp;
    }
    __syncwarp();
  }
#undef FLD
#undef FLDF

  for (int o = 16; o > 0; o >>= 1) {
    rays += __shfl_down_sync(0xffffffffu, rays, o);
    samples += __shfl_down_sync(0xffffffffu, samples, o);
    mism += __shfl_down_sync(0xffffffffu, mism, o);
    if (STATS) {
      tc.box_tests += __shfl_down_sync(0xffffffffu, tc.box_tests, o);
      tc.tri_tests += __shfl_down_sync(0xffffffffu, tc.tri_tests, o);
    }
  }
  if (lane == 0) {
    atomicAdd(&A.counters->rays, rays);
    atomicAdd(&A.counters->samples, samples);
    if (mism) atomicAdd(&A.counters->mismatches, (unsigned long long)mism);
    if (STATS) {
      atomicAdd(&A.counters->box_tests, tc.box_tests);
      atomicAdd(&A.counters->tri_tests, tc.tri_tests);
    }
  }
}

// ---- closest hit of caller-supplied rays (traversal on its own) ---------------------------------------------------------
template <int TRAV, bool SMEM, bool STATS>
__global__ void __launch_bounds__(kBlock) k_trace_rays(const __grid_constant__ KernelArgs A, const float *rays, long long n,
                                                       int *tri_out, float *k_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  size_t used;
  SceneView S = stage_scene<SMEM>(A, smem, &used);
  LaneStack st;
  st.base = reinterpret_cast<float2 *>(smem + used) + threadIdx.x;
  st.stride = kBlock;
  TraceCounters tc;
  tc.box_tests = 0; tc.tri_tests = 0;
  unsigned int mism = 0;
  for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
    const float *r = rays + 6 * i;
    Hit h = closest_hit<TRAV, SMEM, STATS>(S, mk3(r[0], r[1], r[2]), mk3(r[3], r[4], r[5]), st, &tc, &mism);
    tri_out[i] = h.tri;
    k_out[i] = h.k;
  }
  if (STATS) {
    atomicAdd(&A.counters->box_tests, tc.box_tests);
    atomicAdd(&A.counters->tri_tests, tc.tri_tests);
  }
  if (mism) atomicAdd(&A.counters->mismatches, (unsigned long long)mism);
}

// ---- accumulate / clamp / tonemap ---------------------------------------------------------------------------------------------
// out = clamp(sum / spp)   (Raytracing.cl:211-219), float4-vectorised when aligned
__global__ void k_finalize(const float *__restrict__ sums, float *__restrict__ out, long long n, float spp) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = clamp01(sums[i] / spp);
}

// fused multi-GPU reduce + finalize: parts[] may live on peer GPUs (NVLink P2P loads); summed in rank order
struct PartList {
  const float *p[16];
  int n;
};
__global__ void k_reduce_finalize(const __grid_constant__ PartList parts, float *__restrict__ out, long long n4, long long n,
                                  float spp) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = i; v < n4; v += stride) {
    float4 a = reinterpret_cast<const float4 *>(parts.p[0])[v];
    for (int r = 1; r < parts.n; ++r) {
      float4 b = reinterpret_cast<const float4 *>(parts.p[r])[v];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    a.x = clamp01(a.x / spp);
    a.y = clamp01(a.y / spp);
    a.z = clamp01(a.z / spp);
    a.w = clamp01(a.w / spp);
    reinterpret_cast<float4 *>(out)[v] = a;
  }
  for (long long e = 4 * n4 + i; e < n; e += stride) {
    float a = parts.p[0][e];
    for (int r = 1; r < parts.n; ++r) a += parts.p[r][e];
    out[e] = clamp01(a / spp);
  }
}

// ImgProcessing.cl:1-9
__global__ void k_img_processing(const float *__restrict__ in, float *__restrict__ out, long long n, long long global) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < global; i += stride) {
    if (i < n) {
      float p = in[i];
      p = (1.0f < p) ? 1.0f : p;  // OpenCL min(x, y) = y < x ? y : x  (a NaN input stays NaN)
      out[i] = cr_pow(p, 2.2f);
    }
  }
}

// ---- probes -----------------------------------------------------------------------------------------------------------------------
__global__ void k_math_probe(int fn, const float *a, const float *b, long long n, float *out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = a[i], y = b[i], r = 0.0f;
  switch (fn) {
    case 0: r = cr_sin(x); break;
    case 1: r = cr_cos(x); break;
    case 2: r = cr_acos(x); break;
    case 3: r = cr_asin(x); break;
    case 4: r = cr_atan2(x, y); break;
    case 5: r = cr_tan(x); break;
    case 6: r = cr_pow(x, y); break;
    case 7: r = div_safe(y) ? div_by(x, y, __frcp_rn(y)) : __fdiv_rn(x, y); break;
    case 8: r = sqrtf(x); break;
    case 9: { float s, c; cr_sincos(x, &s, &c); r = s; break; }
    case 10: { float s, c; cr_sincos(x, &s, &c); r = c; break; }
  }
  out[i] = r;
}

__global__ void k_philox_probe(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
  uint32_t o[4];
  philox4x32_10(c0, c1, c2, c3, k0, k1, o);
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2]; out[3] = o[3];
}

}  // namespace b200rt
