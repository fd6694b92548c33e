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
// Lane phases.  A lane always holds at most one ray; `T.active` says its traversal is still running.
//   PH_NONE    between samples (or no pixel): start the next sample from the cached primary hit
//   PH_SHADE   the segment hit a non-emissive surface: sample the next direction (Raytracing.cl:51-87)
//   PH_BOUNCE  a bounce ray is being traced / has just finished (:82-110)
//   PH_SUN     the sun shadow ray is being traced / has just finished (:115-137)
enum { PH_NONE = 0, PH_SHADE = 1, PH_BOUNCE = 2, PH_SUN = 3 };

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

  // lane state
  int pix = -1, s = 0, j = 0, phase = PH_NONE;
  v3 seg_o = mk3(0, 0, 0), seg_d = mk3(0, 0, 0);  // current segment (R_cam); after an escape: the escaped direction
  float seg_k = 0.0f;
  int seg_tri = -1;
  int seg_type = 0;                                // material type of the surface the pending ray left
  v3 acc = mk3(0, 0, 0), sum = mk3(0, 0, 0);
  rng_state g;
  g.a = 0;
  Trav T;                                          // the pending ray and its traversal
  T.active = false;
  T.best.tri = -1; T.best.k = 1000.0f; T.best_rank = 0; T.cur = 0; T.sp = 0;
  T.R.o = mk3(0, 0, 0); T.R.d = mk3(0, 0, 0); T.R.r = mk3(0, 0, 0); T.R.fast = false;
  bool exhausted = false;

  unsigned long long rays = 0, samples = 0;
  unsigned int mism = 0;
  TraceCounters tc;
  tc.box_tests = 0; tc.tri_tests = 0;

  for (;;) {
    // ======== phase A: every lane that is not traversing is moved forward until it has a ray again ==========
    bool new_ray = false;

    // A1 resolve a finished trace
    if (pix >= 0 && !T.active && (phase == PH_BOUNCE || phase == PH_SUN)) {
      const Hit h = T.best;
      bool end_sample = false;
      if (phase == PH_BOUNCE) {
        if (h.tri >= 0) {
          int mat = __float_as_int(ld4<SMEM>(S.tris + 3 * (size_t)h.tri + 2).y);
          Material mb = load_material(S.mats, mat);
          seg_o = T.R.o; seg_d = T.R.d; seg_k = h.k; seg_tri = h.tri;
          if (mb.type != 0) {
            if (j == F.max_bounce) {  // :99-103
              acc = mk3(0.0f, 0.0f, 0.0f);
              end_sample = true;
            } else {
              ++j;
              phase = PH_SHADE;
            }
          } else {  // :105-109
            acc = acc * mb.roughness;
            end_sample = true;
          }
        } else {  // escaped: shadow ray towards the sun from the same origin (:115-124)
          seg_d = T.R.d;
          T.R.d = F.sun_dir;
          phase = PH_SUN;
          new_ray = true;
        }
      } else {  // PH_SUN, :125-137
        v3 sun = mk3(0.0f, 0.0f, 0.0f);
        if (h.tri < 0) {
          if (seg_type != 3) sun = mk3(F.sun_power, F.sun_power, F.sun_power);
        } else {
          int mat = __float_as_int(ld4<SMEM>(S.tris + 3 * (size_t)h.tri + 2).y);
          Material ms = load_material(S.mats, mat);
          if (ms.type == 3) sun = ms.color * F.sun_power;
        }
        v3 envl = ibl_lookup(F, A.ibl, seg_d) * F.ibl_power;
        acc = acc * (sun + envl);
        end_sample = true;
      }
      if (end_sample) {
        sum = sum + acc;  // :207
        ++s;
        ++samples;
        phase = PH_NONE;
        if (s >= F.s1) {
          write_pixel(A, pix, sum);
          pix = -1;
        }
      }
    }

    // A2 refill: lanes without a pixel fetch the next work item (one warp-aggregated atomic)
    {
      unsigned int need = __ballot_sync(0xffffffffu, pix < 0);
      if (need != 0u && !exhausted) {
        unsigned int base = 0;
        const int leader = __ffs(need) - 1;
        const unsigned int cnt = __popc(need);
        if ((int)lane == leader) base = atomicAdd(A.work_counter, cnt);
        base = __shfl_sync(0xffffffffu, base, leader);
        if (base + cnt >= (unsigned)A.n_work) exhausted = true;
        if (pix < 0) {
          unsigned int w = base + __popc(need & lt_mask);
          if (w < (unsigned)A.n_work) {
            int i = work_to_pixel(A, w);
            if (i >= 0 && A.prim_tri[i] >= 0) {
              pix = i;
              s = F.s0;
              sum = mk3(0.0f, 0.0f, 0.0f);
              g.a = (uint32_t)i;  // Raytracing.cl:171 (imgSize receives imgDim, so seed0 = i)
              phase = PH_NONE;
            }
          }
        }
      }
    }

    // A3 start of a sample: reload the cached primary hit (Raytracing.cl:195-201)
    if (pix >= 0 && phase == PH_NONE) {
      float4 dk = A.prim_dirk[pix];
      seg_o = F.cam_pos;
      seg_d = mk3(dk.x, dk.y, dk.z);
      seg_k = dk.w;
      seg_tri = A.prim_tri[pix];
      acc = mk3(1.0f, 1.0f, 1.0f);
      j = 0;
      phase = PH_SHADE;
    }

    // A4 shade: choose the next direction and fold BRDF * cos / pdf into the sample (:51-87)
    if (pix >= 0 && phase == PH_SHADE) {
      float4 t2 = ld4<SMEM>(S.tris + 3 * (size_t)seg_tri + 2);
      float4 nn = ld4<SMEM>(S.normals + seg_tri);
      v3 n = mk3(nn.x, nn.y, nn.z);
      Material m = load_material(S.mats, __float_as_int(t2.y));
      v3 nd, brdf;
      float inv_pdf;
      if (m.type == 3) {
        nd = seg_d;
        brdf = m.color;
        inv_pdf = 1.0f / fabsf(dot(nd, unit(n)));
      } else {
        float u0, u1;
        if (F.rng_mode == 0) draw2<0>(g, (uint32_t)pix, (uint32_t)s, (uint32_t)j, F.key0, F.key1, &u0, &u1);
        else draw2<1>(g, (uint32_t)pix, (uint32_t)s, (uint32_t)j, F.key0, F.key1, &u0, &u1);
        if (m.type == 1) {
          nd = sample_cosine(n, u0, u1, &inv_pdf);
          brdf = m.color * (1.0f / 3.14f);
        } else {
          nd = sample_uniform(n, u0, u1, &inv_pdf);
          brdf = bsdf_ggx(m, neg3(seg_d), nd, n);
        }
      }
      T.R.o = seg_o + unit(seg_d) * seg_k;  // :79 — no offset along the normal
      T.R.d = nd;
      float att = inv_pdf * fabsf(dot(nd, unit(n)));
      acc = (acc * brdf) * att;
      seg_type = m.type;
      phase = PH_BOUNCE;
      new_ray = true;
    }

    // A5 the new ray enters the tree
    if (new_ray) {
      rays++;
      if (TRAV == 0) {
        trav_begin<SMEM, STATS>(S, T, T.R.o, T.R.d, &tc);
      } else {  // reference / verify traversal: the whole walk at once
        T.best = closest_hit<TRAV, SMEM, STATS>(S, T.R.o, T.R.d, st, &tc, &mism);
        T.active = false;
      }
    }

    // ======== phase B: advance all running traversals, one node per turn, while enough lanes take part ==========
    const unsigned int alive = exhausted ? __ballot_sync(0xffffffffu, pix >= 0) : 0xffffffffu;
    if (alive == 0u) break;
    if (TRAV == 0) {
      const int quorum = min(A.quorum, __popc(alive));
      for (;;) {
        const unsigned int act = __ballot_sync(0xffffffffu, T.active);
        if (__popc(act) < quorum) break;
        if (T.active) trav_step<SMEM, STATS>(S, T, st, &tc);
      }
    }
  }

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
