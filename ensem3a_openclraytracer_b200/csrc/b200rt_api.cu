// b200rt — host side of the C ABI declared in include/b200rt.h.
//
// Owns the CUDA context objects of one GPU, validates and repacks the reference-layout scene
// buffers into the float4 layouts rt_trace.cuh consumes, evaluates the per-frame
// constants, and launches the kernels of rt_kernels.cuh.  No CPU rendering path exists here: every
// entry point that produces pixels, hits or tonemapped values does so by launching a kernel.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/b200rt.h"
#include "rt_kernels.cuh"
#include "scene_repack.h"

using namespace b200rt;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool borrowed = false;  // a helper context's view of its parent's scene buffer: never freed or resized here
};

}  // namespace

struct b200rt_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;      // stream in use
  cudaStream_t own_stream = nullptr;  // created by b200rt_create
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> kev;  // per-launch events of the last time_kernels render
  int kev_used = 0;
  std::string err;
  int sm_count = 0;
  size_t smem_optin = 0;

  // scene
  bool have_scene = false;
  bool have_scene_cached = false;  // scene_hash / ibl_hash are valid
  uint64_t scene_hash = 0, mat_hash = 0;
  DevBuf d_nodes, d_tris, d_ctris, d_normals, d_tboxes, d_frames, d_mats, d_bvh9, d_leafcnt;
  int n_nodes9 = 0, n_inner = 0, n_tris = 0, n_mats = 0;
  int depth = 0, ref_stack_need = 0;
  int cull_depth = 0;      // depth of the tree the node records hold (the culling tree, scene_repack.h)
  bool canonical = true;
  int root_ref = 0;
  int node_f4 = 2;
  float grid_base[3] = {0, 0, 0}, grid_pitch[3] = {1, 1, 1};
  uint32_t root_w[3] = {0x7fff0000u, 0x7fff0000u, 0x7fff0000u};
  float cull_abs = 0.0f, cmax = 0.0f;
  int fast_ok = 1;
  // k_trace tuning knobs (B200RT_QUORUM / B200RT_REFILL_MIN / B200RT_TRI_QUORUM / B200RT_STEPS override)
  int quorum = 12, refill_min = 8, tri_quorum = 2;
  int carveout = -1;       // B200RT_CARVEOUT: k_trace's shared-memory carve-out (-1 driver's choice, 0 computed, else per cent)
  int max_trace_ctas = 0;  // B200RT_MAX_TRACE_CTAS: cap on resident k_trace CTAs per SM (0 = what fits)
  int stream_trace_ctas = 4;  // B200RT_STREAM_TRACE_CTAS: the cap while three or more sample streams share the GPU (0 = none)
  int frame_trace_cap = 0; // the cap of the frame being enqueued (set by render_frame)
  int compact_every = 4;   // B200RT_COMPACT_EVERY: wavefront iterations between two compactions of the path list
  std::vector<int32_t> tri_mat;  // for re-validating material edits
  DevBuf d_light;                // triangles whose material is emissive, ascending (opt-in light sampling)
  int n_light = 0;
  DevBuf d_pS, d_pL, d_pR;       // light sampling: three more float4 of path state

  // environment map
  bool have_ibl = false;
  uint64_t ibl_hash = 0;
  cudaArray_t ibl_array = nullptr;
  cudaTextureObject_t ibl_tex = 0;
  int ibl_w = 0, ibl_h = 0;

  // per-frame
  DevBuf d_prim_dirk, d_prim_tri, d_out, d_misc, d_tmp_a, d_tmp_b;
  DevBuf d_pA, d_pB, d_pC, d_pHit, d_list0, d_list1, d_cnt;  // wavefront path state (rt_kernels.cuh)
  DevBuf d_slots, d_part_count;                               // order-preserving list compaction
  DeviceCounters *d_counters = nullptr;

  b200rt_stats stats;
  bool stats_pending = false;  // async render enqueued; counters/events not read yet

  // sample streams (b200rt_opts.sample_streams): helper contexts on the same GPU, each with its own stream and path
  // state, that borrow this context's scene and environment
  std::vector<b200rt_ctx *> helpers;
  bool is_helper = false;
  uint64_t scene_epoch = 1, shared_epoch = 0;   // helpers re-borrow when the parent's scene / materials / map changed
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_done = nullptr;
  DevBuf d_part;               // this context's own partial sums when the frame is cut into streams
  int streams_used = 1;        // of the last render
};

namespace {

// B200RT_CULL_TREE=0 keeps the caller's tree topology in the node records (development A/B; default: own tree)
bool own_cull_tree() {
  const char *q = getenv("B200RT_CULL_TREE");
  return !(q && q[0] == '0');
}

int fail(b200rt_ctx *c, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(c, B200RT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

}  // namespace

static_assert(kRefStack == kRefStackMax, "scene_repack.h and rt_trace.cuh disagree on the exact walk's stack");
size_t b200rt::lane_smem_bytes_host(int stack_depth) { return lane_smem_bytes(stack_depth); }

namespace {

int ensure(b200rt_ctx *c, DevBuf &b, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (b.borrowed) return fail(c, B200RT_ERR_INVALID, "internal: resize of a borrowed scene buffer");
  if (b.cap >= bytes) return 0;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  CU(cudaMalloc(&b.p, bytes));
  b.cap = bytes;
  return 0;
}

uint64_t hash_bytes(const void *p, size_t n, uint64_t h) {
  const unsigned char *b = static_cast<const unsigned char *>(p);
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t w;
    memcpy(&w, b + i, 8);
    h = (h ^ w) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
  }
  for (; i < n; ++i) h = (h ^ b[i]) * 0x100000001B3ull;
  return h ^ (uint64_t)n;
}

// correctly-rounded binary32 cos/sin/tan on the host: binary64 libm, one rounding (rt_math.cuh contract)
float h_cos(float a) { return (float)std::cos((double)a); }
float h_sin(float a) { return (float)std::sin((double)a); }
float h_tan(float a) { return (float)std::tan((double)a); }

rotor h_rotor(float angle, v3 axis) {
  float half = angle * 0.5f;
  return make_rotor(h_cos(half), h_sin(half), axis);
}

const float kDeg2Rad = 3.14f / 180.0f;

// Per-frame constants.  Raytracing.cl:24,27,33-35,115-118 and MathLib.cl:73-74.
void frame_setup(const b200rt_ctx *c, const float *cam, const float *env, int width, int height, int spp,
                 int max_bounce, const b200rt_opts &o, FrameParams *F) {
  F->width = width;
  F->height = height;
  F->cam_pos = mk3(cam[0], cam[1], cam[2]);
  F->focal_off = 1.0f / (2.0f * h_tan(cam[9] / 2.0f));
  F->step = (float)(1.0 / (double)cam[6]);
  F->cam_rx = h_rotor(cam[3] * kDeg2Rad, mk3(1, 0, 0));
  F->cam_ry = h_rotor(cam[4] * kDeg2Rad, mk3(0, 1, 0));
  F->cam_rz = h_rotor(cam[5] * kDeg2Rad, mk3(0, 0, 1));
  v3 s = mk3(1, 1, 1);
  if (env) {
    s = apply_rotor(h_rotor(env[0] * kDeg2Rad, mk3(1, 0, 0)), s);
    s = apply_rotor(h_rotor(env[1] * kDeg2Rad, mk3(0, 1, 0)), s);
    s = apply_rotor(h_rotor(env[2] * kDeg2Rad, mk3(0, 0, 1)), s);
  }
  F->sun_dir = s;
  F->sun_power = env ? env[3] : 0.0f;
  F->ibl_power = env ? env[4] : 0.0f;
  F->ibl_r1 = h_rotor(90 * kDeg2Rad, mk3(1, 0, 0));
  F->ibl_r2 = h_rotor(90 * kDeg2Rad, mk3(0, 1, 0));
  F->ibl_w = c->ibl_w;
  F->ibl_h = c->ibl_h;
  F->spp = spp;
  F->max_bounce = max_bounce;
  F->s0 = o.sample_begin;
  F->s1 = o.sample_end;
  if (F->s1 <= 0) { F->s0 = 0; F->s1 = spp; }
  F->pixel_begin = o.pixel_begin;
  F->pixel_end = o.pixel_end;
  if (F->pixel_end <= 0) { F->pixel_begin = 0; F->pixel_end = width * height; }
  F->tile_row_mod = o.tile_row_mod;
  F->tile_row_rem = o.tile_row_rem;
  F->out_mode = o.output;
  F->rng_mode = o.rng_mode;
  F->key0 = (uint32_t)(o.seed & 0xffffffffull);
  F->key1 = (uint32_t)(o.seed >> 32);
  F->sampling = o.sampling;
}

size_t scene_smem_bytes(const b200rt_ctx *c) { return (size_t)c->n_inner * 48 + (size_t)c->n_tris * 48; }

bool use_smem_scene(const b200rt_ctx *c) { return c->node_f4 == 3; }

// entries of the per-lane shared-memory traversal stack: the near-first walk holds at most one entry per level;
// trees walked in reference order keep their stack in thread-local memory instead
int stack_entries(const b200rt_ctx *c) { return c->canonical ? c->cull_depth + 2 : 2; }

size_t smem_bytes(const b200rt_ctx *c, bool smem_scene) {
  return lane_smem_bytes(stack_entries(c)) + (smem_scene ? scene_smem_bytes(c) : 0);
}

int effective_stack_cap(const b200rt_ctx *c, const b200rt_opts &o);

void fill_args(b200rt_ctx *c, const FrameParams &F, const b200rt_opts &o, float *d_out, KernelArgs *A) {
  A->F = F;
  SceneView &S = A->S;
  S.nodes = static_cast<const uint4 *>(c->d_nodes.p);
  S.node_f4 = c->node_f4;
  S.tris = static_cast<const float4 *>(c->d_tris.p);
  S.ctris = static_cast<const float4 *>(c->d_ctris.p);
  S.normals = static_cast<const float4 *>(c->d_normals.p);
  S.tboxes = static_cast<const float4 *>(c->d_tboxes.p);
  S.frames = static_cast<const float4 *>(c->d_frames.p);
  S.mats = static_cast<const float *>(c->d_mats.p);
  S.bvh9 = static_cast<const float *>(c->d_bvh9.p);
  S.leafcnt = static_cast<const int *>(c->d_leafcnt.p);
  S.root_ref = c->root_ref;
  for (int i = 0; i < 3; ++i) {
    S.grid_base[i] = c->grid_base[i];
    S.grid_pitch[i] = c->grid_pitch[i];
    S.root_w[i] = c->root_w[i];
  }
  S.cull_abs = c->cull_abs;
  S.cmax = c->cmax;
  {
    // traversal-stack entry layout (rt_trace.cuh, LaneStack): 2^E above every culling limit, refs in the low bits
    const int E = std::ilogb((1001.0f + c->cull_abs) * 1.01f) + 1;
    S.st_bias = (uint32_t)(E - 32 + 127) << 23;
    unsigned long long m = 1;
    while (m < (unsigned long long)std::max(1, c->n_inner) * (unsigned long long)std::max(1, c->node_f4)) m <<= 1;
    S.st_rmask = (uint32_t)(m - 1ull);
  }
  S.stack_cap = effective_stack_cap(c, o);
  S.fast_ok = c->fast_ok;
  A->ibl = c->ibl_tex;
  A->prim_dirk = static_cast<float4 *>(c->d_prim_dirk.p);
  A->prim_tri = static_cast<int *>(c->d_prim_tri.p);
  A->out = d_out;
  A->pA = static_cast<float4 *>(c->d_pA.p);
  A->pB = static_cast<float4 *>(c->d_pB.p);
  A->pC = static_cast<float4 *>(c->d_pC.p);
  A->pHit = static_cast<int2 *>(c->d_pHit.p);
  A->light = static_cast<const int *>(c->d_light.p);
  A->n_light = c->n_light;
  A->pS = static_cast<float4 *>(c->d_pS.p);
  A->pL = static_cast<float4 *>(c->d_pL.p);
  A->pR = static_cast<float4 *>(c->d_pR.p);
  A->list[0] = static_cast<int *>(c->d_list0.p);
  A->list[1] = static_cast<int *>(c->d_list1.p);
  A->cnt = static_cast<unsigned int *>(c->d_cnt.p);
  A->slots = static_cast<int *>(c->d_slots.p);
  A->part_count = static_cast<unsigned int *>(c->d_part_count.p);
  A->wc = nullptr;
  A->compact_every = c->compact_every;
  A->counters = c->d_counters;
  A->n_nodes = c->n_inner;
  A->n_tris = c->n_tris;
  A->stack_depth = stack_entries(c);
  A->tiles_x = (F.width + 7) / 8;
  int tiles_y = (F.height + 3) / 4;
  A->n_work = A->tiles_x * tiles_y * 32;
  A->quorum = c->quorum;
  A->refill_min = c->refill_min;
  A->tri_quorum = c->tri_quorum;
  A->validate = 0;
}

// The fast path needs a strict two-child tree with nested boxes; any other tree is walked in reference order.  When
// the caller asked for FAST (a walk that never drops a push) the substitute keeps that promise as far as its
// thread-local stack reaches: it runs with the largest cap instead of the reference's 20.
int effective_traversal(const b200rt_ctx *c, const b200rt_opts &o) {
  if (!c->canonical) return B200RT_TRAVERSAL_REFERENCE;
  return o.traversal;
}

int effective_stack_cap(const b200rt_ctx *c, const b200rt_opts &o) {
  if (!c->canonical && o.traversal == B200RT_TRAVERSAL_FAST) return kRefStack;
  return o.stack_cap <= 0 ? 20 : (o.stack_cap > kRefStack ? kRefStack : o.stack_cap);
}

template <typename K>
int set_smem_attr(b200rt_ctx *c, K kernel, size_t bytes) {
  if (bytes > 48 * 1024) CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <typename K>
int persistent_grid(b200rt_ctx *c, K kernel, size_t smem, int *grid) {
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, smem));
  if (per_sm < 1) return fail(c, B200RT_ERR_CUDA, "kernel does not fit on an SM (smem %zu)", smem);
  *grid = per_sm * c->sm_count;
  return 0;
}

template <int TRAV, bool SMEM, bool PARITY, bool STATS>
int launch_primary_t(b200rt_ctx *c, const KernelArgs &A, int *tri_out, float *k_out) {
  auto k = k_primary<TRAV, SMEM, PARITY, STATS>;
  size_t smem = smem_bytes(c, SMEM);
  int grid = 0;
  if (set_smem_attr(c, k, smem)) return B200RT_ERR_CUDA;
  if (persistent_grid(c, k, smem, &grid)) return B200RT_ERR_CUDA;
  int need = (A.n_work + kBlock - 1) / kBlock;
  if (grid > need) grid = need;
  k<<<grid, kBlock, smem, c->stream>>>(A, tri_out, k_out);
  CU(cudaGetLastError());
  c->stats.kernel_launches++;
  return 0;
}

template <int TRAV, bool SMEM>
int launch_primary_ts(b200rt_ctx *c, const KernelArgs &A, bool parity, bool stats, int *tri_out, float *k_out) {
  if (TRAV == 2) stats = false;
  if (parity) return stats ? launch_primary_t<TRAV, SMEM, true, (TRAV != 2)>(c, A, tri_out, k_out)
                           : launch_primary_t<TRAV, SMEM, true, false>(c, A, tri_out, k_out);
  return stats ? launch_primary_t<TRAV, SMEM, false, (TRAV != 2)>(c, A, tri_out, k_out)
               : launch_primary_t<TRAV, SMEM, false, false>(c, A, tri_out, k_out);
}

int launch_primary(b200rt_ctx *c, const KernelArgs &A, int trav, bool smem, bool parity, bool stats, int *tri_out,
                   float *k_out) {
  switch (trav * 2 + (smem ? 1 : 0)) {
    case 0: return launch_primary_ts<0, false>(c, A, parity, stats, tri_out, k_out);
    case 1: return launch_primary_ts<0, true>(c, A, parity, stats, tri_out, k_out);
    case 2: return launch_primary_ts<1, false>(c, A, parity, stats, tri_out, k_out);
    case 3: return launch_primary_ts<1, true>(c, A, parity, stats, tri_out, k_out);
    case 4: return launch_primary_ts<2, false>(c, A, parity, stats, tri_out, k_out);
    case 5: return launch_primary_ts<2, true>(c, A, parity, stats, tri_out, k_out);
  }
  return fail(c, B200RT_ERR_INVALID, "bad traversal mode %d", trav);
}

// Launch geometry of the wavefront kernels, resolved once per frame (the occupancy queries are host calls)
struct WaveLaunch {
  void (*trace)(const KernelArgs, int) = nullptr;
  int trace_grid = 0;
  size_t trace_smem = 0;
  int shade_grid = 0;
};

template <int TRAV, bool SMEM, bool STATS>
int prepare_trace_t(b200rt_ctx *c, WaveLaunch *w) {
  auto k = k_trace<TRAV, SMEM, STATS>;
  w->trace_smem = smem_bytes(c, SMEM);
  if (set_smem_attr(c, k, w->trace_smem)) return B200RT_ERR_CUDA;
  if (persistent_grid(c, k, w->trace_smem, &w->trace_grid)) return B200RT_ERR_CUDA;
  if (c->carveout >= 0) {
    // shared memory is carved out of the SM's L1: ask for what the resident CTAs need (+1 KB each, the system's share)
    // and no more, so that the rest keeps caching nodes and triangles
    int pct = c->carveout;
    if (pct == 0) {
      const size_t need = (size_t)(w->trace_grid / c->sm_count) * (w->trace_smem + 1024);
      pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
    }
    if (pct > 100) pct = 100;
    CU(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
  }
  // With several sample streams on the GPU a k_trace that fills every SM's register file leaves the other streams'
  // kernels nothing but its tail to run in; capped, one stream's shading and another's tracing are resident side by side
  // (bench scene -4 %, Cornell 1080p -12 %, measured with 3 streams and 4 CTAs per SM).
  const int cap = c->max_trace_ctas > 0 ? c->max_trace_ctas : c->frame_trace_cap;
  if (cap > 0 && w->trace_grid > cap * c->sm_count) w->trace_grid = cap * c->sm_count;
  w->trace = k;
  return 0;
}

int prepare_wave(b200rt_ctx *c, int trav, bool smem, bool stats, WaveLaunch *w) {
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_shade<false>, kShadeBlock, 0));
  if (per_sm < 1) return fail(c, B200RT_ERR_CUDA, "k_shade does not fit on an SM");
  w->shade_grid = per_sm * c->sm_count;
  int key = trav * 4 + (smem ? 2 : 0) + (stats ? 1 : 0);
  switch (key) {
    case 0: return prepare_trace_t<0, false, false>(c, w);
    case 1: return prepare_trace_t<0, false, true>(c, w);
    case 2: return prepare_trace_t<0, true, false>(c, w);
    case 3: return prepare_trace_t<0, true, true>(c, w);
    case 4: return prepare_trace_t<1, false, false>(c, w);
    case 5: return prepare_trace_t<1, false, true>(c, w);
    case 6: return prepare_trace_t<1, true, false>(c, w);
    case 7: return prepare_trace_t<1, true, true>(c, w);
    case 8: case 9: return prepare_trace_t<2, false, false>(c, w);
    case 10: case 11: return prepare_trace_t<2, true, false>(c, w);
  }
  return fail(c, B200RT_ERR_INVALID, "bad traversal mode %d", trav);
}

template <int TRAV, bool SMEM, bool STATS>
int launch_trace_t(b200rt_ctx *c, const KernelArgs &A, const float *rays, long long n, int *tri, float *k) {
  auto kern = k_trace_rays<TRAV, SMEM, STATS>;
  size_t smem = smem_bytes(c, SMEM);
  if (set_smem_attr(c, kern, smem)) return B200RT_ERR_CUDA;
  int grid = 0;
  if (persistent_grid(c, kern, smem, &grid)) return B200RT_ERR_CUDA;
  long long need = (n + kBlock - 1) / kBlock;
  if (grid > need) grid = (int)(need < 1 ? 1 : need);
  kern<<<grid, kBlock, smem, c->stream>>>(A, rays, n, tri, k);
  CU(cudaGetLastError());
  c->stats.kernel_launches++;
  return 0;
}

int read_counters_one(b200rt_ctx *c) {
  DeviceCounters h;
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(&h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  c->stats.rays = h.rays;
  c->stats.primary_rays = h.primary_rays;
  c->stats.box_tests = h.box_tests;
  c->stats.tri_tests = h.tri_tests;
  c->stats.mismatches = h.mismatches;
  c->stats.samples = h.samples;
  c->stats.revalidated = (int32_t)(h.revalidated > 0x7fffffffull ? 0x7fffffffull : h.revalidated);
  c->stats.exact_walks = (int32_t)(h.exact_walks > 0x7fffffffull ? 0x7fffffffull : h.exact_walks);
  float ms = 0;
  if (cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]) == cudaSuccess) c->stats.primary_ms = ms;
  if (cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]) == cudaSuccess) c->stats.trace_ms = ms;
  if (cudaEventElapsedTime(&ms, c->ev[0], c->ev[2]) == cudaSuccess) c->stats.total_ms = ms;
  c->stats.shade_kernel_ms = 0.0f;
  c->stats.trace_kernel_ms = 0.0f;
  // B200RT_DUMP_WAVE=<file> (development): one line per wavefront iteration — live paths, k_shade ms, k_trace ms
  FILE *dump = nullptr;
  std::vector<unsigned int> live;
  if (c->kev_used > 1)
    if (const char *path = getenv("B200RT_DUMP_WAVE")) {
      dump = fopen(path, "w");
      live.resize((size_t)c->kev_used / 2 + 2, 0u);
      if (cudaMemcpy(live.data(), c->d_cnt.p, live.size() * sizeof(unsigned int), cudaMemcpyDeviceToHost) != cudaSuccess) live.assign(live.size(), 0u);
    }
  float shade_ms = 0.0f;
  for (int i = 0; i + 1 < c->kev_used; ++i)  // event i sits before launch i; launches alternate shade, trace, shade, ...
    if (cudaEventElapsedTime(&ms, c->kev[i], c->kev[i + 1]) == cudaSuccess) {
      ((i & 1) ? c->stats.trace_kernel_ms : c->stats.shade_kernel_ms) += ms;
      if (dump) {
        if (i & 1) fprintf(dump, "%d %u %.4f %.4f\n", i / 2, live[(size_t)i / 2], shade_ms, ms);
        else shade_ms = ms;
      }
    }
  if (dump) fclose(dump);
  c->kev_used = 0;
  c->stats_pending = false;
  return 0;
}

// Counters and timings of the last render.  A frame cut into sample streams sums the work of its helper contexts
// (the primary rays, which every stream traces for itself, count once) and is timed from the fork to the reduce.
int read_counters(b200rt_ctx *c) {
  int rc = read_counters_one(c);
  if (rc || c->streams_used <= 1) return rc;
  for (int k = 1; k < c->streams_used; ++k) {
    b200rt_ctx *h = c->helpers[(size_t)k - 1];
    rc = read_counters_one(h);
    if (rc) return fail(c, rc, "sample stream %d: %s", k, h->err.c_str());
    c->stats.rays += h->stats.rays - h->stats.primary_rays;
    c->stats.box_tests += h->stats.box_tests;
    c->stats.tri_tests += h->stats.tri_tests;
    c->stats.mismatches += h->stats.mismatches;
    c->stats.samples += h->stats.samples;
    c->stats.revalidated += h->stats.revalidated;
    c->stats.exact_walks += h->stats.exact_walks;
    c->stats.kernel_launches += h->stats.kernel_launches;
    c->stats.shade_kernel_ms += h->stats.shade_kernel_ms;   // kernels of different streams overlap: these are sums of
    c->stats.trace_kernel_ms += h->stats.trace_kernel_ms;   // durations, not a partition of total_ms
    c->stats.primary_ms = std::max(c->stats.primary_ms, h->stats.primary_ms);
  }
  float ms = 0;
  if (cudaEventElapsedTime(&ms, c->ev_fork, c->ev_done) == cudaSuccess) {
    c->stats.total_ms = ms;
    c->stats.trace_ms = ms - c->stats.primary_ms;
  }
  c->stats.kernel_launches += 1;  // the reduce
  return 0;
}

int check_frame_args(b200rt_ctx *c, const float *cam, int width, int height, int spp, int max_bounce,
                     const b200rt_opts &o, bool need_env, const float *env) {
  if (!cam) return fail(c, B200RT_ERR_INVALID, "cam is NULL");
  if (need_env && !env) return fail(c, B200RT_ERR_INVALID, "envData is NULL");
  if (!c->have_scene) return fail(c, B200RT_ERR_NO_SCENE, "no scene: call b200rt_set_scene first");
  if (width <= 0 || height <= 0 || (long long)width * height > 0x7fffffffLL / 4)
    return fail(c, B200RT_ERR_INVALID, "bad frame size %d x %d", width, height);
  if ((int)cam[6] != width)
    return fail(c, B200RT_ERR_INVALID, "cam[6] = %g but width = %d (the kernel derives rows/columns from cam[6])",
                (double)cam[6], width);
  if (spp <= 0) return fail(c, B200RT_ERR_INVALID, "spp must be positive (got %d)", spp);
  if (max_bounce < 0) return fail(c, B200RT_ERR_INVALID, "maxBounce must be >= 0 (got %d)", max_bounce);
  // the per-path state packs the sample index into 16 bits and the bounce into 8 (rt_kernels.cuh meta_pack)
  if (spp > 65535) return fail(c, B200RT_ERR_UNSUPPORTED, "spp %d exceeds 65535 per launch; accumulate several launches with sample ranges", spp);
  if (max_bounce > 254) return fail(c, B200RT_ERR_UNSUPPORTED, "maxBounce %d exceeds 254", max_bounce);
  if (o.rng_mode != B200RT_RNG_REFERENCE && o.rng_mode != B200RT_RNG_PHILOX)
    return fail(c, B200RT_ERR_INVALID, "bad rng_mode %d", o.rng_mode);
  if (o.traversal < 0 || o.traversal > 2) return fail(c, B200RT_ERR_INVALID, "bad traversal %d", o.traversal);
  if (o.sampling < 0 || o.sampling > (B200RT_SAMPLING_IMPORTANCE | B200RT_SAMPLING_LIGHTS))
    return fail(c, B200RT_ERR_INVALID, "bad sampling mode %d", o.sampling);
  if (o.output != B200RT_OUT_FINAL && o.output != B200RT_OUT_SUMS)
    return fail(c, B200RT_ERR_INVALID, "bad output mode %d", o.output);
  if (o.sample_end > 0) {
    if (o.sample_begin < 0 || o.sample_begin >= o.sample_end || o.sample_end > spp)
      return fail(c, B200RT_ERR_INVALID, "bad sample range [%d,%d) for spp %d", o.sample_begin, o.sample_end, spp);
    if (o.sample_begin != 0 && o.rng_mode == B200RT_RNG_REFERENCE)
      return fail(c, B200RT_ERR_INVALID,
                  "a sample range that does not start at 0 needs B200RT_RNG_PHILOX: the reference generator is one "
                  "serial stream per pixel");
    if ((o.sample_begin != 0 || o.sample_end != spp) && o.output != B200RT_OUT_SUMS)
      return fail(c, B200RT_ERR_INVALID, "a partial sample range only makes sense with B200RT_OUT_SUMS");
  }
  if (o.tile_row_mod > 1 && (o.tile_row_rem < 0 || o.tile_row_rem >= o.tile_row_mod))
    return fail(c, B200RT_ERR_INVALID, "bad tile row split %d mod %d", o.tile_row_rem, o.tile_row_mod);
  if (o.pixel_end > 0 && (o.pixel_begin < 0 || o.pixel_begin >= o.pixel_end || o.pixel_end > width * height))
    return fail(c, B200RT_ERR_INVALID, "bad pixel range [%d,%d)", o.pixel_begin, o.pixel_end);
  return 0;
}

int render_impl(b200rt_ctx *c, const float *cam, const float *env, int width, int height, int spp, int max_bounce,
                const b200rt_opts *opts, float *d_out) {
  b200rt_opts o;
  if (opts) o = *opts; else b200rt_default_opts(&o);
  c->streams_used = 1;
  int rc = check_frame_args(c, cam, width, height, spp, max_bounce, o, true, env);
  if (rc) return rc;
  if (!c->have_ibl) return fail(c, B200RT_ERR_NO_SCENE, "no environment map: call b200rt_set_ibl first");
  CU(cudaSetDevice(c->device));
  const size_t npix = (size_t)width * height;
  FrameParams F;
  frame_setup(c, cam, env, width, height, spp, max_bounce, o, &F);
  // every live path traces exactly one ray per iteration and a sample needs at most max_bounce + 1 bounce rays
  // plus one sun ray (Raytracing.cl:46-137)
  // with light sampling every surface that scatters adds one shadow ray towards an emitter: up to 2 (max_bounce + 1) + 1
  const bool nee = (o.sampling & B200RT_SAMPLING_LIGHTS) != 0 && c->n_light > 0;
  const long long per_sample = nee ? 2 * ((long long)max_bounce + 1) + 1 : (long long)max_bounce + 2;
  const long long n_iter_ll = (long long)(F.s1 - F.s0) * per_sample;
  if (n_iter_ll > (1 << 22)) return fail(c, B200RT_ERR_UNSUPPORTED, "%lld wavefront iterations exceed 2^22", n_iter_ll);
  const int n_iter = (int)n_iter_ll;
  if (ensure(c, c->d_prim_dirk, npix * sizeof(float4))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_prim_tri, npix * sizeof(int))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_pA, npix * sizeof(float4))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_pB, npix * sizeof(float4))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_pC, npix * sizeof(float4))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_pHit, npix * sizeof(int2))) return B200RT_ERR_CUDA;
  if (nee) {
    if (ensure(c, c->d_pS, npix * sizeof(float4))) return B200RT_ERR_CUDA;
    if (ensure(c, c->d_pL, npix * sizeof(float4))) return B200RT_ERR_CUDA;
    if (ensure(c, c->d_pR, npix * sizeof(float4))) return B200RT_ERR_CUDA;
  }
  if (ensure(c, c->d_list0, npix * sizeof(int))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_list1, npix * sizeof(int))) return B200RT_ERR_CUDA;
  const size_t n_work = (size_t)((width + 7) / 8) * ((height + 3) / 4) * 32;   // tile-ordered work items of k_primary
  if (ensure(c, c->d_slots, n_work * sizeof(int))) return B200RT_ERR_CUDA;
  const size_t n_cnt = (size_t)n_iter + 2;             // cnt[0 .. n_iter + 1]
  const size_t cnt_bytes = (n_cnt + (size_t)n_iter + 1) * sizeof(unsigned int);  // then wc[0 .. n_iter]
  if (ensure(c, c->d_cnt, cnt_bytes)) return B200RT_ERR_CUDA;
  KernelArgs A;
  fill_args(c, F, o, d_out, &A);
  A.wc = A.cnt + n_cnt;
  const int trav = effective_traversal(c, o);
  A.validate = (trav == B200RT_TRAVERSAL_FAST) ? 1 : 0;
  const bool smem = use_smem_scene(c);
  WaveLaunch wl;
  rc = prepare_wave(c, trav, smem, o.collect_stats != 0, &wl);
  if (rc) return rc;
  if (ensure(c, c->d_part_count, (size_t)wl.shade_grid * sizeof(unsigned int))) return B200RT_ERR_CUDA;
  A.part_count = static_cast<unsigned int *>(c->d_part_count.p);
  A.slots = static_cast<int *>(c->d_slots.p);
  c->stats.kernel_launches = 0;
  c->stats.sample_streams = 1;
  c->stats.scene_in_smem = smem ? 1 : 0;
  CU(cudaMemsetAsync(c->d_counters, 0, sizeof(DeviceCounters), c->stream));
  CU(cudaMemsetAsync(c->d_cnt.p, 0, cnt_bytes, c->stream));
  CU(cudaEventRecord(c->ev[0], c->stream));
  rc = launch_primary(c, A, trav, smem, false, o.collect_stats != 0, nullptr, nullptr);
  if (rc) return rc;
  // the first path list: k_primary's live pixels in tile order
  k_count_parts<<<wl.shade_grid, kCompactBlock, 0, c->stream>>>(A.slots, (unsigned)A.n_work, A.part_count);
  k_compact<<<wl.shade_grid, kCompactBlock, 0, c->stream>>>(A.slots, nullptr, (unsigned)A.n_work, A.part_count, A.list[0], A.cnt);
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev[1], c->stream));
  c->kev_used = 0;
  if (o.time_kernels) {
    const size_t need = 2 * (size_t)n_iter + 2;
    while (c->kev.size() < need) {
      cudaEvent_t e;
      CU(cudaEventCreate(&e));
      c->kev.push_back(e);
    }
  }
  int n_compactions = 0;
  for (int it = 0; it <= n_iter; ++it) {
    if (o.time_kernels) CU(cudaEventRecord(c->kev[c->kev_used++], c->stream));
    if (nee) k_shade<true><<<wl.shade_grid, kShadeBlock, 0, c->stream>>>(A, it);
    else k_shade<false><<<wl.shade_grid, kShadeBlock, 0, c->stream>>>(A, it);
    if (it < n_iter && (it + 1) % A.compact_every == 0) {  // close the gaps the finished pixels left, order kept (timed with k_shade)
      k_compact<<<wl.shade_grid, kCompactBlock, 0, c->stream>>>(A.slots, A.cnt + it, 0u, A.part_count, A.list[(it + 1) & 1], A.cnt + it + 1);
      ++n_compactions;
    }
    if (o.time_kernels) CU(cudaEventRecord(c->kev[c->kev_used++], c->stream));
    if (it < n_iter) wl.trace<<<wl.trace_grid, kBlock, wl.trace_smem, c->stream>>>(A, it);
  }
  CU(cudaGetLastError());
  // k_count_parts + k_compact for the first list; k_shade + k_trace per iteration; the compactions; the closing k_shade
  c->stats.kernel_launches += 2 + 2 * n_iter + n_compactions + 1;
  c->stats.wave_iterations = n_iter;
  CU(cudaEventRecord(c->ev[2], c->stream));
  c->stats_pending = true;
  return 0;
}

// ---- sample streams ------------------------------------------------------------------------------------------------
// One path per pixel is in flight inside a render (a pixel's samples run in order).  With the counter-based
// generator a frame's sample range can be cut into N contiguous parts that do not depend on each other; each part is
// an ordinary render into its own partial-sum buffer, issued on its own CUDA stream by a helper context that borrows
// this context's scene, and one k_reduce_finalize on this context's stream adds the parts in order.  The GPU then
// holds N samples of a pixel at once: small frames fill it, and on large frames one part's shading overlaps another
// part's tracing (the two kernels are bound by different units).
void borrow_scene(b200rt_ctx *p, b200rt_ctx *h) {
  auto lend = [](DevBuf &dst, const DevBuf &src) { dst.p = src.p; dst.cap = src.cap; dst.borrowed = true; };
  lend(h->d_nodes, p->d_nodes); lend(h->d_tris, p->d_tris); lend(h->d_ctris, p->d_ctris); lend(h->d_normals, p->d_normals); lend(h->d_tboxes, p->d_tboxes);
  lend(h->d_frames, p->d_frames); lend(h->d_mats, p->d_mats); lend(h->d_bvh9, p->d_bvh9); lend(h->d_leafcnt, p->d_leafcnt);
  lend(h->d_light, p->d_light);
  h->n_light = p->n_light;
  h->n_nodes9 = p->n_nodes9; h->n_inner = p->n_inner; h->n_tris = p->n_tris; h->n_mats = p->n_mats;
  h->node_f4 = p->node_f4; h->depth = p->depth; h->cull_depth = p->cull_depth; h->ref_stack_need = p->ref_stack_need; h->canonical = p->canonical;
  h->root_ref = p->root_ref;
  for (int k = 0; k < 3; ++k) {
    h->grid_base[k] = p->grid_base[k]; h->grid_pitch[k] = p->grid_pitch[k];
    h->root_w[k] = p->root_w[k];
  }
  h->cull_abs = p->cull_abs; h->cmax = p->cmax; h->fast_ok = p->fast_ok;
  h->quorum = p->quorum; h->refill_min = p->refill_min; h->tri_quorum = p->tri_quorum;
  h->max_trace_ctas = p->max_trace_ctas; h->compact_every = p->compact_every; h->carveout = p->carveout;
  h->have_scene = p->have_scene;
  h->ibl_tex = p->ibl_tex; h->ibl_w = p->ibl_w; h->ibl_h = p->ibl_h; h->have_ibl = p->have_ibl;
  h->shared_epoch = p->scene_epoch;
}

int resolve_streams(const b200rt_opts &o, int width, int height, int s0, int s1) {
  if (o.rng_mode != B200RT_RNG_PHILOX) return 1;   // the reference generator is one serial stream per pixel
  int n = o.sample_streams;
  if (n < 0) {
    const long long npix = (long long)width * height;
    n = npix >= 1000000 ? 3 : (npix >= 200000 ? 4 : 8);
  }
  if (n > 16) n = 16;
  if (n > s1 - s0) n = s1 - s0;
  return n < 1 ? 1 : n;
}

int render_frame(b200rt_ctx *c, const float *cam, const float *env, int width, int height, int spp, int max_bounce,
                 const b200rt_opts *opts, float *d_out) {
  b200rt_opts o;
  if (opts) o = *opts; else b200rt_default_opts(&o);
  if (o.sample_streams < -1 || o.sample_streams > 16)
    return fail(c, B200RT_ERR_INVALID, "bad sample_streams %d (-1 automatic, 0 / 1 off, up to 16)", o.sample_streams);
  int s0 = o.sample_begin, s1 = o.sample_end;
  if (s1 <= 0) { s0 = 0; s1 = spp; }
  const int n = (width > 0 && height > 0 && spp > 0) ? resolve_streams(o, width, height, s0, s1) : 1;
  if (!c->is_helper) c->frame_trace_cap = 0;
  if (n <= 1 || c->is_helper) return render_impl(c, cam, env, width, height, spp, max_bounce, &o, d_out);
  int rc = check_frame_args(c, cam, width, height, spp, max_bounce, o, true, env);
  if (rc) return rc;
  if (!c->have_ibl) return fail(c, B200RT_ERR_NO_SCENE, "no environment map: call b200rt_set_ibl first");
  CU(cudaSetDevice(c->device));
  while ((int)c->helpers.size() < n - 1) {
    b200rt_ctx *h = nullptr;
    rc = b200rt_create(c->device, &h);
    if (rc) return fail(c, rc, "sample stream context: %s", b200rt_last_error(nullptr));
    h->is_helper = true;
    c->helpers.push_back(h);
  }
  c->frame_trace_cap = n >= 3 ? c->stream_trace_ctas : 0;
  for (int k = 1; k < n; ++k) c->helpers[(size_t)k - 1]->frame_trace_cap = c->frame_trace_cap;
  if (!c->ev_fork) {
    CU(cudaEventCreate(&c->ev_fork));
    CU(cudaEventCreate(&c->ev_done));
  }
  const size_t bytes = (size_t)width * height * 3 * sizeof(float);
  if (ensure(c, c->d_part, bytes)) return B200RT_ERR_CUDA;
  CU(cudaEventRecord(c->ev_fork, c->stream));   // the parts start after whatever the caller enqueued before this frame
  std::vector<int> rcs((size_t)n, 0);
  const int total = s1 - s0, base = total / n, extra = total % n;
#pragma omp parallel for num_threads(n) schedule(static, 1)
  for (int k = 0; k < n; ++k) {
    b200rt_ctx *h = k == 0 ? c : c->helpers[(size_t)k - 1];
    b200rt_opts ok = o;
    ok.sample_streams = 1;
    ok.output = B200RT_OUT_SUMS;
    ok.sample_begin = s0 + k * base + (k < extra ? k : extra);
    ok.sample_end = ok.sample_begin + base + (k < extra ? 1 : 0);
    int r = 0;
    if (cudaSetDevice(c->device) != cudaSuccess) r = B200RT_ERR_CUDA;
    float *part = nullptr;
    if (!r && k > 0) {
      if (h->shared_epoch != c->scene_epoch) borrow_scene(c, h);
      if (!h->ev_join && cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) r = B200RT_ERR_CUDA;
      if (!r && cudaStreamWaitEvent(h->stream, c->ev_fork, 0) != cudaSuccess) r = B200RT_ERR_CUDA;
      if (!r) r = ensure(h, h->d_part, bytes);
      part = static_cast<float *>(h->d_part.p);
    } else if (!r) {
      part = static_cast<float *>(c->d_part.p);
    }
    if (!r && cudaMemsetAsync(part, 0, bytes, h->stream) != cudaSuccess) r = B200RT_ERR_CUDA;
    if (!r) r = render_impl(h, cam, env, width, height, spp, max_bounce, &ok, part);
    if (!r && k > 0 && cudaEventRecord(h->ev_join, h->stream) != cudaSuccess) r = B200RT_ERR_CUDA;
    if (r == B200RT_ERR_CUDA && h->err.empty()) h->err = "CUDA call failed while enqueueing a sample stream";
    rcs[(size_t)k] = r;
  }
  for (int k = 0; k < n; ++k)
    if (rcs[(size_t)k]) {
      b200rt_ctx *h = k == 0 ? c : c->helpers[(size_t)k - 1];
      const std::string msg = h->err;
      return fail(c, rcs[(size_t)k], "sample stream %d: %s", k, msg.c_str());
    }
  // join on this context's stream, add the parts in order; a whole frame is finalised, a partial range stays sums
  PartList pl;
  pl.n = n;
  pl.p[0] = static_cast<const float *>(c->d_part.p);
  for (int k = 1; k < n; ++k) {
    CU(cudaStreamWaitEvent(c->stream, c->helpers[(size_t)k - 1]->ev_join, 0));
    pl.p[k] = static_cast<const float *>(c->helpers[(size_t)k - 1]->d_part.p);
  }
  const long long nn = (long long)width * height * 3;
  const bool aligned = (reinterpret_cast<uintptr_t>(d_out) % 16) == 0;
  const int grid = (int)std::min<long long>((nn / 4 + 255) / 256 + 1, (long long)c->sm_count * 8);
  k_reduce_finalize<<<grid, 256, 0, c->stream>>>(pl, d_out, aligned ? nn / 4 : 0, nn, o.output == B200RT_OUT_FINAL ? (float)spp : 0.0f);
  CU(cudaGetLastError());
  CU(cudaEventRecord(c->ev_done, c->stream));
  c->streams_used = n;
  c->stats.sample_streams = n;
  c->stats_pending = true;
  return 0;
}

}  // namespace

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

const char *b200rt_version(void) { return "b200rt 0.2 (sm_100a)"; }

int b200rt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void b200rt_default_opts(b200rt_opts *o) {
  memset(o, 0, sizeof *o);
  o->rng_mode = B200RT_RNG_REFERENCE;
  o->traversal = B200RT_TRAVERSAL_FAST;
  o->stack_cap = 20;
  o->output = B200RT_OUT_FINAL;
}

const char *b200rt_last_error(const b200rt_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int b200rt_create(int device, b200rt_ctx **out) {
  b200rt_ctx *c = nullptr;
  if (!out) return fail(nullptr, B200RT_ERR_INVALID, "out_ctx is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, B200RT_ERR_CUDA, "no CUDA device available (%s); b200rt has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= n) return fail(nullptr, B200RT_ERR_INVALID, "device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return fail(nullptr, B200RT_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, B200RT_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library contains sm_100a code only", device,
                prop.major, prop.minor);
  c = new b200rt_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  memset(&c->stats, 0, sizeof c->stats);
  auto env_int = [](const char *name, int lo, int hi, int *dst) {
    if (const char *q = getenv(name)) {
      int v = atoi(q);
      if (v >= lo && v <= hi) *dst = v;
    }
  };
  env_int("B200RT_QUORUM", 1, 32, &c->quorum);
  env_int("B200RT_REFILL_MIN", 1, 32, &c->refill_min);
  env_int("B200RT_TRI_QUORUM", 1, 32, &c->tri_quorum);
  env_int("B200RT_MAX_TRACE_CTAS", 1, 32, &c->max_trace_ctas);
  env_int("B200RT_STREAM_TRACE_CTAS", 0, 32, &c->stream_trace_ctas);
  env_int("B200RT_CARVEOUT", -1, 100, &c->carveout);
  env_int("B200RT_COMPACT_EVERY", 1, 1 << 20, &c->compact_every);
  auto bail = [&](const char *what, cudaError_t err) {
    fail(nullptr, B200RT_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
    delete c;
    return B200RT_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  c->stream = c->own_stream;
  for (int i = 0; i < 3; ++i)
    if ((e = cudaEventCreate(&c->ev[i])) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaMalloc(&c->d_counters, sizeof(DeviceCounters))) != cudaSuccess) return bail("cudaMalloc", e);
  *out = c;
  return 0;
}

void b200rt_destroy(b200rt_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (b200rt_ctx *h : c->helpers) b200rt_destroy(h);
  c->helpers.clear();
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->d_part.p) cudaFree(c->d_part.p);
  if (c->is_helper) {  // the environment map belongs to the parent
    c->ibl_tex = 0;
    c->ibl_array = nullptr;
  }
  DevBuf *bufs[] = {&c->d_nodes, &c->d_tris, &c->d_ctris, &c->d_normals, &c->d_tboxes, &c->d_frames, &c->d_mats, &c->d_bvh9, &c->d_leafcnt, &c->d_prim_dirk,
                    &c->d_prim_tri, &c->d_out, &c->d_misc, &c->d_tmp_a, &c->d_tmp_b, &c->d_pA, &c->d_pB, &c->d_pC,
                    &c->d_pHit, &c->d_list0, &c->d_list1, &c->d_cnt, &c->d_slots, &c->d_part_count, &c->d_light, &c->d_pS,
                    &c->d_pL, &c->d_pR};
  for (DevBuf *b : bufs)
    if (b->p && !b->borrowed) cudaFree(b->p);
  if (c->ibl_tex) cudaDestroyTextureObject(c->ibl_tex);
  if (c->ibl_array) cudaFreeArray(c->ibl_array);
  if (c->d_counters) cudaFree(c->d_counters);
  for (int i = 0; i < 3; ++i)
    if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (cudaEvent_t e : c->kev) cudaEventDestroy(e);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

int b200rt_set_materials(b200rt_ctx *c, const float *mat, int64_t n_mat) {
  if (!c) return B200RT_ERR_INVALID;
  if (!mat || n_mat <= 0 || n_mat % 6) return fail(c, B200RT_ERR_INVALID, "material_data must hold 6 floats per material (got %lld)", (long long)n_mat);
  const int nm = (int)(n_mat / 6);
  for (int m = 0; m < nm; ++m) {
    float t = mat[6 * m];
    if (!(t >= 0.0f && t < 4.0f))
      return fail(c, B200RT_ERR_INVALID, "material %d has type %g; the kernel defines types 0..3 only (Raytracing.cl:58-78)", m, (double)t);
  }
  for (int32_t m : c->tri_mat)
    if (m < 0 || m >= nm) return fail(c, B200RT_ERR_INVALID, "a triangle uses material %d but only %d materials were given", m, nm);
  CU(cudaSetDevice(c->device));
  uint64_t h = hash_bytes(mat, (size_t)n_mat * 4, 0x6d617473ull);
  if (c->n_mats == nm && c->mat_hash != 0 && h == c->mat_hash && c->d_mats.p) return 0;
  if (ensure(c, c->d_mats, (size_t)n_mat * 4)) return B200RT_ERR_CUDA;
  CU(cudaMemcpyAsync(c->d_mats.p, mat, (size_t)n_mat * 4, cudaMemcpyHostToDevice, c->stream));
  // The emitter list of the opt-in light sampling: what FileManager.py:235-240 builds as lightData — the triangles
  // whose material is emissive, in triangle order — derived here from the materials just given, so that it can never
  // be stale after a material edit (the reference's kernel receives lightData and never reads it, Raytracing.cl:163).
  {
    std::vector<int32_t> lights;
    for (size_t t = 0; t < c->tri_mat.size(); ++t)
      if ((int)mat[6 * (size_t)c->tri_mat[t]] == 0) lights.push_back((int32_t)t);
    c->n_light = (int)lights.size();
    if (!lights.empty()) {
      if (ensure(c, c->d_light, lights.size() * sizeof(int32_t))) return B200RT_ERR_CUDA;
      CU(cudaMemcpy(c->d_light.p, lights.data(), lights.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
  }
  CU(cudaStreamSynchronize(c->stream));
  c->n_mats = nm;
  c->mat_hash = h;
  ++c->scene_epoch;
  return 0;
}

}  // extern "C"

namespace {

struct SceneArgs {
  const float *vp; int64_t n_vp;
  const float *vn; int64_t n_vn;
  const int32_t *face; int64_t n_face;
  const float *mat; int64_t n_mat;
  const float *bvh; int64_t n_bvh;
};

// shape checks of the reference-layout buffers; the message goes to the context (or to the create-error slot)
int check_scene_args(b200rt_ctx *c, const SceneArgs &a) {
  if (!a.vp || !a.vn || !a.face || !a.mat || !a.bvh) return fail(c, B200RT_ERR_INVALID, "NULL scene buffer");
  if (a.n_vp <= 0 || a.n_vp % 3) return fail(c, B200RT_ERR_INVALID, "vertex_p length %lld is not a positive multiple of 3", (long long)a.n_vp);
  if (a.n_vn <= 0 || a.n_vn % 3) return fail(c, B200RT_ERR_INVALID, "vertex_n length %lld is not a positive multiple of 3", (long long)a.n_vn);
  if (a.n_face <= 0 || a.n_face % 10) return fail(c, B200RT_ERR_INVALID, "face_data length %lld is not a positive multiple of 10", (long long)a.n_face);
  if (a.n_bvh <= 0 || a.n_bvh % 9) return fail(c, B200RT_ERR_INVALID, "BVH length %lld is not a positive multiple of 9", (long long)a.n_bvh);
  if (a.n_bvh / 9 >= (1 << 24)) return fail(c, B200RT_ERR_UNSUPPORTED, "%lld nodes: float32-encoded child indices are exact only below 2^24 (BVH.py:165)", (long long)(a.n_bvh / 9));
  if (a.n_mat <= 0 || a.n_mat % 6) return fail(c, B200RT_ERR_INVALID, "material_data must hold 6 floats per material (got %lld)", (long long)a.n_mat);
  return 0;
}

uint64_t scene_hash_of(const SceneArgs &a) {
  uint64_t h = 0x7363656e65ull;
  h = hash_bytes(a.vp, (size_t)a.n_vp * 4, h);
  h = hash_bytes(a.vn, (size_t)a.n_vn * 4, h);
  h = hash_bytes(a.face, (size_t)a.n_face * 4, h);
  h = hash_bytes(a.bvh, (size_t)a.n_bvh * 4, h);
  return h;
}

bool scene_is_cached(const b200rt_ctx *c, uint64_t h) { return c->have_scene && c->have_scene_cached && h == c->scene_hash; }

// Copies a repacked scene to the context's GPU and, only after every copy has succeeded, commits the fields the
// kernels' launch geometry and margins derive from.  `R` is shared by every GPU of a multi-GPU handle.
int upload_scene(b200rt_ctx *c, const Repacked &R, const SceneArgs &a, uint64_t h) {
  // From here on the context describes no scene until the new one is complete: a caller that catches an error and
  // resubmits the previous scene must not hit the content-hash shortcut with half-replaced buffers.
  c->have_scene = false;
  c->have_scene_cached = false;
  c->scene_hash = 0;
  const int n_tris = R.n_tris;
  CU(cudaSetDevice(c->device));
  if (ensure(c, c->d_tris, R.tris.size() * sizeof(float4))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_ctris, R.ctris.size() * sizeof(float4))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_normals, R.normals.size() * sizeof(float4))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_tboxes, R.tboxes.size() * sizeof(float4))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_frames, (size_t)n_tris * kFrameVec * sizeof(float4))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_nodes, R.nodes.size() * sizeof(uint4))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_bvh9, (size_t)a.n_bvh * 4)) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_leafcnt, R.leaf_count.size() * sizeof(int32_t))) return B200RT_ERR_CUDA;
  CU(cudaMemcpyAsync(c->d_leafcnt.p, R.leaf_count.data(), R.leaf_count.size() * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_tris.p, R.tris.data(), R.tris.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_ctris.p, R.ctris.data(), R.ctris.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_normals.p, R.normals.data(), R.normals.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_tboxes.p, R.tboxes.data(), R.tboxes.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  if (!R.nodes.empty())
    CU(cudaMemcpyAsync(c->d_nodes.p, R.nodes.data(), R.nodes.size() * sizeof(uint4), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->d_bvh9.p, a.bvh, (size_t)a.n_bvh * 4, cudaMemcpyHostToDevice, c->stream));
  k_tri_frames<<<(n_tris + 127) / 128, 128, 0, c->stream>>>(static_cast<const float4 *>(c->d_normals.p), n_tris,
                                                           static_cast<float4 *>(c->d_frames.p));
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(c->stream));
  c->n_nodes9 = R.n_nodes9;
  c->n_inner = R.n_inner;
  c->n_tris = n_tris;
  c->node_f4 = R.node_f4;
  c->depth = R.depth;
  c->cull_depth = R.cull_depth;
  c->ref_stack_need = R.ref_stack_need;
  c->canonical = R.canonical;
  c->root_ref = R.root_ref;
  for (int k = 0; k < 3; ++k) {
    c->grid_base[k] = R.grid_base[k];
    c->grid_pitch[k] = R.grid_pitch[k];
    c->root_w[k] = R.root_w[k];
  }
  c->cull_abs = R.cull_abs;
  c->cmax = R.cmax;
  c->fast_ok = R.fast_ok;
  c->tri_mat = R.tri_mat;
  c->mat_hash = 0;
  c->n_mats = 0;
  int rc = b200rt_set_materials(c, a.mat, a.n_mat);
  if (rc) return rc;
  c->have_scene = true;
  c->have_scene_cached = true;
  c->scene_hash = h;
  ++c->scene_epoch;
  c->stats.repack_ms = (float)(R.ms_tris + R.ms_walk + R.ms_nodes);
  c->stats.ref_stack_need = c->ref_stack_need;
  c->stats.nodes = R.n_nodes9;
  c->stats.triangles = n_tris;
  c->stats.bvh_depth = c->depth;
  return 0;
}

}  // namespace

extern "C" {

int b200rt_set_scene(b200rt_ctx *c, const float *vp, int64_t n_vp, const float *vn, int64_t n_vn, const float *vuv,
                     int64_t n_vuv, const int32_t *face, int64_t n_face, const float *mat, int64_t n_mat,
                     const int32_t *light, int64_t n_light, const float *bvh, int64_t n_bvh) {
  (void)vuv; (void)n_vuv; (void)light; (void)n_light;  // fetched / passed but never used by the kernel (MathLib.cl:217, Raytracing.cl:163)
  if (!c) return B200RT_ERR_INVALID;
  auto t0 = std::chrono::steady_clock::now();
  const SceneArgs a{vp, n_vp, vn, n_vn, face, n_face, mat, n_mat, bvh, n_bvh};
  int rc = check_scene_args(c, a);
  if (rc) return rc;
  const uint64_t h = scene_hash_of(a);
  if (scene_is_cached(c, h)) {  // geometry unchanged (the UI rebuilds identical arrays per render, UI.py:98)
    rc = b200rt_set_materials(c, mat, n_mat);
  } else {
    Repacked R;
    std::string msg;
    rc = repack_scene(vp, n_vp, vn, n_vn, face, n_face, n_mat / 6, bvh, n_bvh, &R, &msg, own_cull_tree());
    if (rc) {
      c->have_scene = false;   // as before an upload: the context describes no scene after a rejected one
      c->have_scene_cached = false;
      c->scene_hash = 0;
      return fail(c, rc, "%s", msg.c_str());
    }
    rc = upload_scene(c, R, a, h);
  }
  c->stats.upload_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return rc;
}

int b200rt_set_ibl(b200rt_ctx *c, const uint8_t *rgba, int width, int height) {
  if (!c) return B200RT_ERR_INVALID;
  if (!rgba || width <= 0 || height <= 0) return fail(c, B200RT_ERR_INVALID, "bad environment map (%p, %d x %d)", (const void *)rgba, width, height);
  if (width > 131072 || height > 65536) return fail(c, B200RT_ERR_UNSUPPORTED, "environment map %d x %d exceeds the 2-D texture limits", width, height);
  auto t0 = std::chrono::steady_clock::now();
  uint64_t h = hash_bytes(rgba, (size_t)width * height * 4, 0x69626cull);
  if (c->have_ibl && c->ibl_hash != 0 && h == c->ibl_hash && width == c->ibl_w && height == c->ibl_h) return 0;
  CU(cudaSetDevice(c->device));
  if (c->ibl_array && c->ibl_tex && width == c->ibl_w && height == c->ibl_h) {
    // same shape: refill the existing array (allocating / freeing device memory synchronises the whole device,
    // peers and collectives included)
    c->have_ibl = false;
    CU(cudaMemcpy2DToArrayAsync(c->ibl_array, 0, 0, rgba, (size_t)width * 4, (size_t)width * 4, (size_t)height,
                                cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  } else {
    CU(cudaStreamSynchronize(c->stream));
    if (c->ibl_tex) { cudaDestroyTextureObject(c->ibl_tex); c->ibl_tex = 0; }
    if (c->ibl_array) { cudaFreeArray(c->ibl_array); c->ibl_array = nullptr; }
    c->have_ibl = false;
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    CU(cudaMallocArray(&c->ibl_array, &fmt, (size_t)width, (size_t)height));
    CU(cudaMemcpy2DToArrayAsync(c->ibl_array, 0, 0, rgba, (size_t)width * 4, (size_t)width * 4, (size_t)height,
                                cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    cudaResourceDesc res;
    memset(&res, 0, sizeof res);
    res.resType = cudaResourceTypeArray;
    res.res.array.array = c->ibl_array;
    cudaTextureDesc td;
    memset(&td, 0, sizeof td);
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;  // CLK_ADDRESS_CLAMP_TO_EDGE (Raytracing.cl:179)
    td.filterMode = cudaFilterModePoint;                          // integer coordinates address single texels
    td.readMode = cudaReadModeElementType;                        // raw bytes; the kernel divides by 255 itself
    td.normalizedCoords = 0;                                      // CLK_NORMALIZED_COORDS_FALSE
    CU(cudaCreateTextureObject(&c->ibl_tex, &res, &td, nullptr));
  }
  c->ibl_w = width;
  c->ibl_h = height;
  c->ibl_hash = h;
  ++c->scene_epoch;
  c->have_ibl = true;
  c->stats.upload_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

int b200rt_render_device(b200rt_ctx *c, const float *cam, const float *env, int width, int height, int spp,
                         int max_bounce, const b200rt_opts *opts, float *d_out) {
  if (!c) return B200RT_ERR_INVALID;
  if (!d_out) return fail(c, B200RT_ERR_INVALID, "d_out_rgb is NULL");
  return render_frame(c, cam, env, width, height, spp, max_bounce, opts, d_out);
}

int b200rt_render(b200rt_ctx *c, const float *cam, const float *env, int width, int height, int spp, int max_bounce,
                  const b200rt_opts *opts, float *out) {
  if (!c) return B200RT_ERR_INVALID;
  if (!out) return fail(c, B200RT_ERR_INVALID, "out_rgb is NULL");
  if (width <= 0 || height <= 0) return fail(c, B200RT_ERR_INVALID, "bad frame size %d x %d", width, height);
  const size_t bytes = (size_t)width * height * 3 * sizeof(float);
  CU(cudaSetDevice(c->device));
  if (ensure(c, c->d_out, bytes)) return B200RT_ERR_CUDA;
  CU(cudaMemsetAsync(c->d_out.p, 0, bytes, c->stream));
  int rc = render_frame(c, cam, env, width, height, spp, max_bounce, opts, static_cast<float *>(c->d_out.p));
  if (rc) return rc;
  CU(cudaMemcpyAsync(out, c->d_out.p, bytes, cudaMemcpyDeviceToHost, c->stream));  // blocking read-back, KernelLauncher.py:78
  CU(cudaStreamSynchronize(c->stream));
  return read_counters(c);
}

int b200rt_render_rgb8(b200rt_ctx *c, const float *cam, const float *env, int width, int height, int spp, int max_bounce,
                       const b200rt_opts *opts, uint8_t *out) {
  if (!c) return B200RT_ERR_INVALID;
  if (!out) return fail(c, B200RT_ERR_INVALID, "out_rgb8 is NULL");
  if (width <= 0 || height <= 0) return fail(c, B200RT_ERR_INVALID, "bad frame size %d x %d", width, height);
  if (opts && opts->output != B200RT_OUT_FINAL) return fail(c, B200RT_ERR_INVALID, "8-bit output needs B200RT_OUT_FINAL");
  const size_t n = (size_t)width * height * 3;
  CU(cudaSetDevice(c->device));
  if (ensure(c, c->d_out, n * sizeof(float))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_tmp_a, n)) return B200RT_ERR_CUDA;
  CU(cudaMemsetAsync(c->d_out.p, 0, n * sizeof(float), c->stream));
  int rc = render_frame(c, cam, env, width, height, spp, max_bounce, opts, static_cast<float *>(c->d_out.p));
  if (rc) return rc;
  int grid = (int)std::min<long long>(((long long)n + 255) / 256, (long long)c->sm_count * 16);
  k_quantize8<<<grid, 256, 0, c->stream>>>(static_cast<const float *>(c->d_out.p), static_cast<uint8_t *>(c->d_tmp_a.p), (long long)n);
  CU(cudaGetLastError());
  c->stats.kernel_launches++;
  CU(cudaMemcpyAsync(out, c->d_tmp_a.p, n, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return read_counters(c);
}

int b200rt_sync(b200rt_ctx *c) {
  if (!c) return B200RT_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  if (c->stats_pending) return read_counters(c);
  return 0;
}

int b200rt_set_stream(b200rt_ctx *c, void *cuda_stream) {
  if (!c) return B200RT_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  if (c->stats_pending) read_counters(c);
  c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
  return 0;
}

int b200rt_invalidate(b200rt_ctx *c) {
  if (!c) return B200RT_ERR_INVALID;
  c->scene_hash = 0;
  c->mat_hash = 0;
  c->ibl_hash = 0;
  c->have_scene_cached = false;
  return 0;
}

int b200rt_finalize_device(b200rt_ctx *c, const float *d_sums, float *d_out, int64_t n_pixels, int spp) {
  if (!c) return B200RT_ERR_INVALID;
  if (!d_sums || !d_out || n_pixels <= 0 || spp <= 0) return fail(c, B200RT_ERR_INVALID, "bad finalize arguments");
  CU(cudaSetDevice(c->device));
  long long n = (long long)n_pixels * 3;
  int grid = (int)std::min<long long>((n + 255) / 256, (long long)c->sm_count * 8);
  k_finalize<<<grid, 256, 0, c->stream>>>(d_sums, d_out, n, (float)spp);
  CU(cudaGetLastError());
  return 0;
}

int b200rt_reduce_finalize_device(b200rt_ctx *c, const float *const *d_parts, int n_parts, float *d_out,
                                  int64_t n_pixels, int spp) {
  if (!c) return B200RT_ERR_INVALID;
  if (!d_parts || n_parts < 1 || n_parts > 16 || !d_out || n_pixels <= 0 || spp <= 0)
    return fail(c, B200RT_ERR_INVALID, "bad reduce_finalize arguments (1..16 parts)");
  CU(cudaSetDevice(c->device));
  PartList pl;
  pl.n = n_parts;
  bool aligned = (reinterpret_cast<uintptr_t>(d_out) % 16) == 0;
  for (int i = 0; i < n_parts; ++i) {
    if (!d_parts[i]) return fail(c, B200RT_ERR_INVALID, "part %d is NULL", i);
    pl.p[i] = d_parts[i];
    aligned = aligned && (reinterpret_cast<uintptr_t>(d_parts[i]) % 16) == 0;
  }
  long long n = (long long)n_pixels * 3;
  long long n4 = aligned ? n / 4 : 0;
  int grid = (int)std::min<long long>((n / 4 + 255) / 256 + 1, (long long)c->sm_count * 8);
  k_reduce_finalize<<<grid, 256, 0, c->stream>>>(pl, d_out, n4, n, (float)spp);
  CU(cudaGetLastError());
  return 0;
}

int b200rt_primary_hits(b200rt_ctx *c, const float *cam, int width, int height, const b200rt_opts *opts,
                        int32_t *tri_out, float *k_out) {
  if (!c) return B200RT_ERR_INVALID;
  if (!tri_out || !k_out) return fail(c, B200RT_ERR_INVALID, "NULL output");
  b200rt_opts o;
  if (opts) o = *opts; else b200rt_default_opts(&o);
  int rc = check_frame_args(c, cam, width, height, 1, 0, o, false, nullptr);
  if (rc) return rc;
  CU(cudaSetDevice(c->device));
  const size_t npix = (size_t)width * height;
  if (ensure(c, c->d_tmp_a, npix * sizeof(int))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_tmp_b, npix * sizeof(float))) return B200RT_ERR_CUDA;
  FrameParams F;
  frame_setup(c, cam, nullptr, width, height, 1, 0, o, &F);
  KernelArgs A;
  fill_args(c, F, o, nullptr, &A);
  c->stats.kernel_launches = 0;
  CU(cudaMemsetAsync(c->d_tmp_a.p, 0xff, npix * sizeof(int), c->stream));
  CU(cudaMemsetAsync(c->d_tmp_b.p, 0, npix * sizeof(float), c->stream));
  CU(cudaMemsetAsync(c->d_counters, 0, sizeof(DeviceCounters), c->stream));
  CU(cudaEventRecord(c->ev[0], c->stream));
  rc = launch_primary(c, A, effective_traversal(c, o), use_smem_scene(c), true, o.collect_stats != 0, static_cast<int *>(c->d_tmp_a.p),
                      static_cast<float *>(c->d_tmp_b.p));
  if (rc) return rc;
  CU(cudaEventRecord(c->ev[1], c->stream));
  CU(cudaEventRecord(c->ev[2], c->stream));
  CU(cudaMemcpyAsync(tri_out, c->d_tmp_a.p, npix * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(k_out, c->d_tmp_b.p, npix * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return read_counters(c);
}

int b200rt_trace_rays(b200rt_ctx *c, const float *rays, int64_t n, const b200rt_opts *opts, int32_t *tri_out, float *k_out) {
  if (!c) return B200RT_ERR_INVALID;
  if (!rays || !tri_out || !k_out || n <= 0) return fail(c, B200RT_ERR_INVALID, "bad trace_rays arguments");
  if (!c->have_scene) return fail(c, B200RT_ERR_NO_SCENE, "no scene: call b200rt_set_scene first");
  b200rt_opts o;
  if (opts) o = *opts; else b200rt_default_opts(&o);
  if (o.traversal < 0 || o.traversal > 2) return fail(c, B200RT_ERR_INVALID, "bad traversal %d", o.traversal);
  CU(cudaSetDevice(c->device));
  if (ensure(c, c->d_misc, (size_t)n * 6 * sizeof(float))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_tmp_a, (size_t)n * sizeof(int))) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_tmp_b, (size_t)n * sizeof(float))) return B200RT_ERR_CUDA;
  CU(cudaMemcpyAsync(c->d_misc.p, rays, (size_t)n * 6 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(c->d_counters, 0, sizeof(DeviceCounters), c->stream));
  FrameParams F;
  memset(&F, 0, sizeof F);
  KernelArgs A;
  fill_args(c, F, o, nullptr, &A);
  c->stats.kernel_launches = 0;
  const float *dr = static_cast<const float *>(c->d_misc.p);
  int *dt = static_cast<int *>(c->d_tmp_a.p);
  float *dk = static_cast<float *>(c->d_tmp_b.p);
  CU(cudaEventRecord(c->ev[0], c->stream));
  CU(cudaEventRecord(c->ev[1], c->stream));
  int rc;
  const bool smem = use_smem_scene(c), st = o.collect_stats != 0;
  const int trav = effective_traversal(c, o);
  if (trav == 0) rc = smem ? (st ? launch_trace_t<0, true, true>(c, A, dr, n, dt, dk) : launch_trace_t<0, true, false>(c, A, dr, n, dt, dk))
                           : (st ? launch_trace_t<0, false, true>(c, A, dr, n, dt, dk) : launch_trace_t<0, false, false>(c, A, dr, n, dt, dk));
  else if (trav == 1) rc = smem ? (st ? launch_trace_t<1, true, true>(c, A, dr, n, dt, dk) : launch_trace_t<1, true, false>(c, A, dr, n, dt, dk))
                                : (st ? launch_trace_t<1, false, true>(c, A, dr, n, dt, dk) : launch_trace_t<1, false, false>(c, A, dr, n, dt, dk));
  else rc = smem ? launch_trace_t<2, true, false>(c, A, dr, n, dt, dk) : launch_trace_t<2, false, false>(c, A, dr, n, dt, dk);
  if (rc) return rc;
  CU(cudaEventRecord(c->ev[2], c->stream));
  CU(cudaMemcpyAsync(tri_out, dt, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaMemcpyAsync(k_out, dk, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  rc = read_counters(c);
  c->stats.rays = (uint64_t)n;
  return rc;
}

int b200rt_img_processing(b200rt_ctx *c, const float *src, float *dst, int64_t n, int64_t global) {
  if (!c) return B200RT_ERR_INVALID;
  if (!src || !dst || global <= 0 || n < 0) return fail(c, B200RT_ERR_INVALID, "bad img_processing arguments");
  CU(cudaSetDevice(c->device));
  size_t bytes = (size_t)global * sizeof(float);
  if (ensure(c, c->d_tmp_a, bytes)) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_tmp_b, bytes)) return B200RT_ERR_CUDA;
  CU(cudaMemcpyAsync(c->d_tmp_a.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
  // work-items with i >= n leave dst untouched: start from the caller's dst contents
  CU(cudaMemcpyAsync(c->d_tmp_b.p, dst, bytes, cudaMemcpyHostToDevice, c->stream));
  int grid = (int)std::min<long long>(((long long)global + 255) / 256, (long long)c->sm_count * 16);
  k_img_processing<<<grid, 256, 0, c->stream>>>(static_cast<const float *>(c->d_tmp_a.p), static_cast<float *>(c->d_tmp_b.p),
                                                 (long long)n, (long long)global);
  CU(cudaGetLastError());
  c->stats.kernel_launches = 1;
  CU(cudaMemcpyAsync(dst, c->d_tmp_b.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int b200rt_get_stats(const b200rt_ctx *c, b200rt_stats *s) {
  if (!c || !s) return B200RT_ERR_INVALID;
  if (c->stats_pending) {
    int rc = b200rt_sync(const_cast<b200rt_ctx *>(c));
    if (rc) return rc;
  }
  *s = c->stats;
  return 0;
}

int b200rt_math_probe(b200rt_ctx *c, int fn, const float *a, const float *b, int64_t n, float *out) {
  if (!c) return B200RT_ERR_INVALID;
  if (!a || !out || n <= 0 || fn < 0 || fn > 16) return fail(c, B200RT_ERR_INVALID, "bad math_probe arguments");
  CU(cudaSetDevice(c->device));
  size_t bytes = (size_t)n * sizeof(float);
  if (ensure(c, c->d_tmp_a, bytes)) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_tmp_b, bytes)) return B200RT_ERR_CUDA;
  if (ensure(c, c->d_misc, bytes)) return B200RT_ERR_CUDA;
  CU(cudaMemcpyAsync(c->d_tmp_a.p, a, bytes, cudaMemcpyHostToDevice, c->stream));
  if (b) CU(cudaMemcpyAsync(c->d_tmp_b.p, b, bytes, cudaMemcpyHostToDevice, c->stream));
  else CU(cudaMemsetAsync(c->d_tmp_b.p, 0, bytes, c->stream));
  k_math_probe<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(fn, static_cast<const float *>(c->d_tmp_a.p),
                                                                    static_cast<const float *>(c->d_tmp_b.p), (long long)n,
                                                                    static_cast<float *>(c->d_misc.p));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, c->d_misc.p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int b200rt_philox_probe(b200rt_ctx *c, const uint32_t ctr[4], uint32_t key0, uint32_t key1, uint32_t out[4]) {
  if (!c) return B200RT_ERR_INVALID;
  if (!ctr || !out) return fail(c, B200RT_ERR_INVALID, "NULL argument");
  CU(cudaSetDevice(c->device));
  if (ensure(c, c->d_misc, 16)) return B200RT_ERR_CUDA;
  k_philox_probe<<<1, 1, 0, c->stream>>>(ctr[0], ctr[1], ctr[2], ctr[3], key0, key1, static_cast<uint32_t *>(c->d_misc.p));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out, c->d_misc.p, 16, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int b200rt_alloc(b200rt_ctx *c, int64_t bytes, void **d_ptr_out) {
  if (!c) return B200RT_ERR_INVALID;
  if (bytes <= 0 || !d_ptr_out) return fail(c, B200RT_ERR_INVALID, "bad alloc arguments");
  CU(cudaSetDevice(c->device));
  CU(cudaMalloc(d_ptr_out, (size_t)bytes));
  CU(cudaMemsetAsync(*d_ptr_out, 0, (size_t)bytes, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return 0;
}

int b200rt_free(b200rt_ctx *c, void *d_ptr) {
  if (!c) return B200RT_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaFree(d_ptr));
  return 0;
}

int b200rt_ipc_export(b200rt_ctx *c, const void *d_ptr, uint8_t handle_out[64]) {
  if (!c) return B200RT_ERR_INVALID;
  if (!d_ptr || !handle_out) return fail(c, B200RT_ERR_INVALID, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CU(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
  memcpy(handle_out, &h, 64);
  return 0;
}

int b200rt_ipc_open(b200rt_ctx *c, const uint8_t handle[64], void **d_ptr_out) {
  if (!c) return B200RT_ERR_INVALID;
  if (!handle || !d_ptr_out) return fail(c, B200RT_ERR_INVALID, "NULL argument");
  CU(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CU(cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int b200rt_ipc_close(b200rt_ctx *c, void *d_ptr) {
  if (!c) return B200RT_ERR_INVALID;
  CU(cudaSetDevice(c->device));
  CU(cudaIpcCloseMemHandle(d_ptr));
  return 0;
}

}  // extern "C"

// ================================================================================================
// Multi-GPU handle: one host process, one context + stream per GPU, NVLink peer reads for the reduce
// ================================================================================================
// the frame is divided among the GPUs of one NVSwitch box — by contiguous sample ranges with the
// counter-based generator, by interleaved rows of 8x4-pixel tiles with the reference's serial per-pixel generator —
// every GPU renders raw per-pixel sums into its own buffer, and ONE kernel on the first GPU reads all of them through
// peer mappings (NVLink loads), sums in GPU order, divides by spp and clamps (k_reduce_finalize).  Kernel launches
// are issued by one host thread per GPU: a frame is thousands of small launches, and a single thread feeding eight
// GPUs would be the bottleneck.
struct b200rt_multi {
  std::vector<b200rt_ctx *> ctx;
  std::vector<cudaEvent_t> done;   // per GPU: its partial sums are complete
  std::vector<int> peer_ok;        // GPU 0 can read GPU i's memory directly
  std::vector<DevBuf> staging;     // on GPU 0, for GPUs it cannot read directly
  DevBuf d_final;                  // on GPU 0
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;  // on GPU 0's stream, around wait + reduce
  std::string err;
  b200rt_stats stats;
};

namespace {

int mfail(b200rt_multi *m, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (m) m->err = buf; else g_create_error = buf;
  return code;
}

#define MCU(call)                                                                                    \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return mfail(m, B200RT_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// what GPU `r` of `n` renders (ensem3a_openclraytracer_b200/multigpu.py::rank_work is the same rule)
void multi_share(int r, int n, int spp, int rng_mode, b200rt_opts *o) {
  o->output = B200RT_OUT_SUMS;
  if (rng_mode == B200RT_RNG_PHILOX) {
    const int base = spp / n, extra = spp % n;
    const int s0 = r * base + (r < extra ? r : extra);
    o->sample_begin = s0;
    o->sample_end = s0 + base + (r < extra ? 1 : 0);
    o->tile_row_mod = 0;
    o->tile_row_rem = 0;
  } else {
    o->sample_begin = 0;
    o->sample_end = spp;
    o->tile_row_mod = n > 1 ? n : 0;
    o->tile_row_rem = n > 1 ? r : 0;
  }
}

}  // namespace

extern "C" {

const char *b200rt_multi_last_error(const b200rt_multi *m) { return m ? m->err.c_str() : g_create_error.c_str(); }

void b200rt_multi_destroy(b200rt_multi *m) {
  if (!m) return;
  if (!m->ctx.empty() && m->ctx[0]) {
    cudaSetDevice(m->ctx[0]->device);
    cudaStreamSynchronize(m->ctx[0]->stream);
    for (DevBuf &b : m->staging)
      if (b.p) cudaFree(b.p);
    if (m->d_final.p) cudaFree(m->d_final.p);
    if (m->ev_begin) cudaEventDestroy(m->ev_begin);
    if (m->ev_end) cudaEventDestroy(m->ev_end);
  }
  for (size_t i = 0; i < m->ctx.size(); ++i) {
    if (!m->ctx[i]) continue;
    if (i < m->done.size() && m->done[i]) {
      cudaSetDevice(m->ctx[i]->device);
      cudaEventDestroy(m->done[i]);
    }
    b200rt_destroy(m->ctx[i]);
  }
  delete m;
}

int b200rt_multi_create(const int *devices, int n_devices, b200rt_multi **out) {
  b200rt_multi *m = nullptr;
  if (!out) return mfail(nullptr, B200RT_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (!devices || n_devices < 1 || n_devices > 16) return mfail(nullptr, B200RT_ERR_INVALID, "1..16 devices expected (got %d)", n_devices);
  // A device may be listed several times: every entry gets its own context, buffers and stream, so a frame too small
  // to fill one GPU (one path per pixel in flight) runs as several concurrent sample-range renders on it.
  m = new b200rt_multi();
  memset(&m->stats, 0, sizeof m->stats);
  m->ctx.assign((size_t)n_devices, nullptr);
  m->done.assign((size_t)n_devices, nullptr);
  m->peer_ok.assign((size_t)n_devices, 1);
  m->staging.resize((size_t)n_devices);
  for (int i = 0; i < n_devices; ++i) {
    int rc = b200rt_create(devices[i], &m->ctx[i]);   // leaves its message in the create-error slot
    if (rc) { b200rt_multi_destroy(m); return rc; }
    cudaError_t e = cudaEventCreateWithFlags(&m->done[i], cudaEventDisableTiming);
    if (e != cudaSuccess) { mfail(nullptr, B200RT_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e)); b200rt_multi_destroy(m); return B200RT_ERR_CUDA; }
  }
  cudaSetDevice(devices[0]);
  for (int i = 1; i < n_devices; ++i) {
    if (devices[i] == devices[0]) continue;  // the same GPU: plain loads
    int can = 0;
    cudaDeviceCanAccessPeer(&can, devices[0], devices[i]);
    if (can) {
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[i], 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
      can = (e == cudaSuccess);
      if (!can) cudaGetLastError();
    }
    m->peer_ok[i] = can;
  }
  if (cudaEventCreate(&m->ev_begin) != cudaSuccess || cudaEventCreate(&m->ev_end) != cudaSuccess) {
    mfail(nullptr, B200RT_ERR_CUDA, "cudaEventCreate failed");
    b200rt_multi_destroy(m);
    return B200RT_ERR_CUDA;
  }
  *out = m;
  return 0;
}

int b200rt_multi_device_count(const b200rt_multi *m) { return m ? (int)m->ctx.size() : 0; }

b200rt_ctx *b200rt_multi_context(b200rt_multi *m, int index) {
  if (!m || index < 0 || index >= (int)m->ctx.size()) return nullptr;
  return m->ctx[index];
}

int b200rt_multi_set_scene(b200rt_multi *m, const float *vp, int64_t n_vp, const float *vn, int64_t n_vn, const float *vuv,
                           int64_t n_vuv, const int32_t *face, int64_t n_face, const float *mat, int64_t n_mat,
                           const int32_t *light, int64_t n_light, const float *bvh, int64_t n_bvh) {
  (void)vuv; (void)n_vuv; (void)light; (void)n_light;
  if (!m) return B200RT_ERR_INVALID;
  auto t0 = std::chrono::steady_clock::now();
  const SceneArgs a{vp, n_vp, vn, n_vn, face, n_face, mat, n_mat, bvh, n_bvh};
  int rc = check_scene_args(m->ctx[0], a);
  if (rc) { m->err = m->ctx[0]->err; return rc; }
  const uint64_t h = scene_hash_of(a);
  bool all_cached = true;
  for (b200rt_ctx *c : m->ctx) all_cached = all_cached && scene_is_cached(c, h);
  Repacked R;
  if (!all_cached) {  // validated and repacked once, uploaded to every GPU
    std::string msg;
    rc = repack_scene(vp, n_vp, vn, n_vn, face, n_face, n_mat / 6, bvh, n_bvh, &R, &msg, own_cull_tree());
    if (rc) {
      for (b200rt_ctx *c : m->ctx) { c->have_scene = false; c->have_scene_cached = false; c->scene_hash = 0; }
      return mfail(m, rc, "%s", msg.c_str());
    }
  }
  const int n = (int)m->ctx.size();
  std::vector<int> rcs((size_t)n, 0);
#pragma omp parallel for num_threads(n) schedule(static, 1)
  for (int i = 0; i < n; ++i) {
    b200rt_ctx *c = m->ctx[i];
    rcs[i] = scene_is_cached(c, h) ? b200rt_set_materials(c, mat, n_mat) : upload_scene(c, R, a, h);
  }
  for (int i = 0; i < n; ++i)
    if (rcs[i]) return mfail(m, rcs[i], "GPU %d: %s", m->ctx[i]->device, m->ctx[i]->err.c_str());
  m->stats.upload_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  m->stats.nodes = m->ctx[0]->stats.nodes;
  m->stats.triangles = m->ctx[0]->stats.triangles;
  m->stats.bvh_depth = m->ctx[0]->stats.bvh_depth;
  m->stats.repack_ms = m->ctx[0]->stats.repack_ms;
  m->stats.ref_stack_need = m->ctx[0]->stats.ref_stack_need;
  return 0;
}

int b200rt_multi_set_ibl(b200rt_multi *m, const uint8_t *rgba, int width, int height) {
  if (!m) return B200RT_ERR_INVALID;
  const int n = (int)m->ctx.size();
  std::vector<int> rcs((size_t)n, 0);
#pragma omp parallel for num_threads(n) schedule(static, 1)
  for (int i = 0; i < n; ++i) rcs[i] = b200rt_set_ibl(m->ctx[i], rgba, width, height);
  for (int i = 0; i < n; ++i)
    if (rcs[i]) return mfail(m, rcs[i], "GPU %d: %s", m->ctx[i]->device, m->ctx[i]->err.c_str());
  return 0;
}

int b200rt_multi_invalidate(b200rt_multi *m) {
  if (!m) return B200RT_ERR_INVALID;
  for (b200rt_ctx *c : m->ctx) b200rt_invalidate(c);
  return 0;
}

int b200rt_multi_render(b200rt_multi *m, const float *cam, const float *env, int width, int height, int spp, int max_bounce,
                        const b200rt_opts *opts, float *out) {
  if (!m) return B200RT_ERR_INVALID;
  if (!out) return mfail(m, B200RT_ERR_INVALID, "out_rgb is NULL");
  if (width <= 0 || height <= 0) return mfail(m, B200RT_ERR_INVALID, "bad frame size %d x %d", width, height);
  b200rt_opts o;
  if (opts) o = *opts; else b200rt_default_opts(&o);
  if (o.output != B200RT_OUT_FINAL || o.sample_end > 0 || o.tile_row_mod > 1)
    return mfail(m, B200RT_ERR_INVALID, "the multi-GPU handle divides the frame itself: leave output / sample range / tile rows at their defaults");
  const int n = (int)m->ctx.size();
  const size_t npix = (size_t)width * height, bytes = npix * 3 * sizeof(float);
  b200rt_ctx *c0 = m->ctx[0];
  MCU(cudaSetDevice(c0->device));
  if (m->d_final.cap < bytes) {
    if (m->d_final.p) cudaFree(m->d_final.p);
    m->d_final.p = nullptr; m->d_final.cap = 0;
    MCU(cudaMalloc(&m->d_final.p, bytes));
    m->d_final.cap = bytes;
  }
  // ---- every GPU renders its share into its own buffer; one host thread per GPU issues the launches ------------
  std::vector<int> rcs((size_t)n, 0);
  const int tile_rows = (height + 3) / 4;
#pragma omp parallel for num_threads(n) schedule(static, 1)
  for (int i = 0; i < n; ++i) {
    b200rt_ctx *c = m->ctx[i];
    b200rt_opts oi = o;
    multi_share(i, n, spp, o.rng_mode, &oi);
    int rc = 0;
    if (cudaSetDevice(c->device) != cudaSuccess) rc = B200RT_ERR_CUDA;
    if (!rc) rc = ensure(c, c->d_out, bytes);
    if (!rc && cudaMemsetAsync(c->d_out.p, 0, bytes, c->stream) != cudaSuccess) rc = B200RT_ERR_CUDA;
    const bool empty = (oi.sample_begin == oi.sample_end) || (oi.tile_row_mod > 1 && oi.tile_row_rem >= tile_rows);
    c->stats.rays = 0; c->stats.samples = 0; c->stats.total_ms = 0; c->stats.kernel_launches = 0;
    if (!rc && !empty) rc = render_frame(c, cam, env, width, height, spp, max_bounce, &oi, static_cast<float *>(c->d_out.p));
    if (!rc && cudaEventRecord(m->done[i], c->stream) != cudaSuccess) rc = B200RT_ERR_CUDA;
    if (rc == B200RT_ERR_CUDA && c->err.empty()) c->err = "CUDA call failed while enqueueing the frame";
    rcs[i] = rc;
  }
  for (int i = 0; i < n; ++i)
    if (rcs[i]) return mfail(m, rcs[i], "GPU %d: %s", m->ctx[i]->device, m->ctx[i]->err.c_str());
  // ---- GPU 0: wait for the peers (device-side), then one kernel reads every partial buffer and finalises ----------
  MCU(cudaSetDevice(c0->device));
  MCU(cudaEventRecord(m->ev_begin, c0->stream));
  PartList pl;
  pl.n = n;
  for (int i = 0; i < n; ++i) {
    if (i > 0) MCU(cudaStreamWaitEvent(c0->stream, m->done[i], 0));
    const float *src = static_cast<const float *>(m->ctx[i]->d_out.p);
    if (i > 0 && !m->peer_ok[i]) {  // no direct mapping: stage the partial sums on GPU 0
      DevBuf &sb = m->staging[i];
      if (sb.cap < bytes) {
        if (sb.p) cudaFree(sb.p);
        sb.p = nullptr; sb.cap = 0;
        MCU(cudaMalloc(&sb.p, bytes));
        sb.cap = bytes;
      }
      MCU(cudaMemcpyPeerAsync(sb.p, c0->device, src, m->ctx[i]->device, bytes, c0->stream));
      src = static_cast<const float *>(sb.p);
    }
    pl.p[i] = src;
  }
  const long long nn = (long long)npix * 3;
  const int grid = (int)std::min<long long>((nn / 4 + 255) / 256 + 1, (long long)c0->sm_count * 8);
  k_reduce_finalize<<<grid, 256, 0, c0->stream>>>(pl, static_cast<float *>(m->d_final.p), nn / 4, nn, (float)spp);
  MCU(cudaGetLastError());
  MCU(cudaEventRecord(m->ev_end, c0->stream));
  MCU(cudaMemcpyAsync(out, m->d_final.p, bytes, cudaMemcpyDeviceToHost, c0->stream));
  MCU(cudaStreamSynchronize(c0->stream));
  // ---- statistics: work summed over the GPUs, device time = the slowest GPU + the reduce ---------------------------
  b200rt_stats agg = m->stats;
  const float upload_ms = agg.upload_ms;
  memset(&agg, 0, sizeof agg);
  agg.upload_ms = upload_ms;
  float slowest = 0.0f;
  for (int i = 0; i < n; ++i) {
    b200rt_ctx *c = m->ctx[i];
    if (c->stats_pending) {
      MCU(cudaSetDevice(c->device));
      int rc = read_counters(c);
      if (rc) return mfail(m, rc, "GPU %d: %s", c->device, c->err.c_str());
    }
    agg.rays += c->stats.rays;
    agg.box_tests += c->stats.box_tests;
    agg.tri_tests += c->stats.tri_tests;
    agg.mismatches += c->stats.mismatches;
    agg.samples += c->stats.samples;
    agg.kernel_launches += c->stats.kernel_launches;
    agg.revalidated += c->stats.revalidated;
    agg.exact_walks += c->stats.exact_walks;
    slowest = std::max(slowest, c->stats.total_ms);
    agg.primary_ms = std::max(agg.primary_ms, c->stats.primary_ms);
    agg.trace_ms = std::max(agg.trace_ms, c->stats.trace_ms);
  }
  float red = 0.0f;
  MCU(cudaSetDevice(c0->device));
  cudaEventElapsedTime(&red, m->ev_begin, m->ev_end);  // includes waiting for the slowest peer after GPU 0 finished
  agg.total_ms = std::max(slowest, c0->stats.total_ms + red);
  agg.kernel_launches += 1;
  agg.nodes = c0->stats.nodes;
  agg.triangles = c0->stats.triangles;
  agg.bvh_depth = c0->stats.bvh_depth;
  agg.scene_in_smem = c0->stats.scene_in_smem;
  agg.repack_ms = c0->stats.repack_ms;
  agg.ref_stack_need = c0->stats.ref_stack_need;
  m->stats = agg;
  return 0;
}

int b200rt_multi_get_stats(const b200rt_multi *m, b200rt_stats *s) {
  if (!m || !s) return B200RT_ERR_INVALID;
  *s = m->stats;
  return 0;
}

}  // extern "C"

