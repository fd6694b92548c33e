// TEST INFRASTRUCTURE — not part of the product path.
//
// Host driver around the reference's own OpenCL-C sources.  build_ref.py copies
// /root/reference/Kernels/*.cl into oracle/_ref/ (git-ignored) applying two
// mechanical regex rewrites (vector-literal casts -> constructor calls,
// .yzw/.xyz swizzles -> accessor calls) and this file #includes the result, so
// every arithmetic statement executed below is the reference's own text.
// The only code of ours is the per-pixel loop that plays the role of the
// OpenCL NDRange (KernelLauncher.py:76-77) and ref_primary(), which re-walks
// the reference's traversal loop (MathLib.cl:245-280) with the reference's own
// interNode/makeTri/intersect to report WHICH triangle the primary ray keeps —
// the reference's hitInfo has no triangle-id field (MathLib.cl:23-30).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <omp.h>

#include "cl_shim.h"

namespace clref {
thread_local int g_global_id = 0;
thread_local Counters g_cnt = {0, 0, 0, 0};

#include "Raytracing.cl"     // pulls MathLib.cl -> stack.cl (transpiled copies in oracle/_ref)
#include "ImgProcessing.cl"
}  // namespace clref

using namespace clref;

extern "C" {

// Runs the reference `Raytracing` kernel for work-items [i0, i1).
// Argument order mirrors KernelLauncher.py:76-77.  counters_out (may be NULL):
// {rays, box_tests, tri_tests, rand_calls}, non-zero only in the counting build.
void ref_raytrace(float *out, const float *vp, const float *vn, const float *vuv, const int *face,
                  const int *light, const float *mat, const float *bvh, const float *cam,
                  const float *env, int triCount, int lightCount, int imgSize, int spp, int maxBounce,
                  int iblW, int iblH, const unsigned char *rgba, int i0, int i1, int nthreads,
                  unsigned long long *counters_out) {
  image2d_t ibl = {iblW, iblH, rgba};
  unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel reduction(+ : c0, c1, c2, c3)
  {
    g_cnt = Counters{0, 0, 0, 0};
#pragma omp for schedule(dynamic, 64)
    for (int i = i0; i < i1; ++i) {
      g_global_id = i;
      Raytracing(out, (float *)vp, (float *)vn, (float *)vuv, (int *)face, (int *)light, (float *)mat,
                 (float *)bvh, (float *)cam, (float *)env, triCount, lightCount, imgSize, spp,
                 maxBounce, ibl);
    }
    c0 += g_cnt.rays;
    c1 += g_cnt.box_tests;
    c2 += g_cnt.tri_tests;
    c3 += g_cnt.rand_calls;
  }
  if (counters_out) {
    counters_out[0] = c0;
    counters_out[1] = c1;
    counters_out[2] = c2;
    counters_out[3] = c3;
  }
}

// Primary ray of work-items [i0,i1): direction (3 floats/pixel), hit flag, hit
// distance k and material as returned by the reference's rayTrace(), plus the
// kept triangle id found by re-walking the same loop (-1 on a miss).
void ref_primary(const float *vp, const float *vn, const float *vuv, const int *face, const float *bvh,
                 const float *cam, int triCount, int i0, int i1, float *dir_out, int *hit_out,
                 float *k_out, int *mat_out, int *tri_out, float *n_out) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = i0; i < i1; ++i) {
    ray r = genCameraRay(i, (float *)cam);
    hitInfo H = rayTrace(r, (float *)vp, (float *)vn, (float *)vuv, (int *)face, triCount, (float *)bvh);
    // same visiting order and acceptance rule as MathLib.cl:245-280
    int kept = -1;
    float bestK = 1000.0f;
    Stack S;
    S.capacity = 20;
    S.top = -1;
    push(&S, 0);
    while (!isEmpty(&S)) {
      int curr = (int)pop(&S);
      if (interNode(r, (float *)bvh, curr)) {
        int t = (int)(bvh[9 * curr + 8]);
        if (t != -1) {
          tri T = makeTri(t, (float *)vp, (float *)vn, (float *)vuv, (int *)face, triCount);
          hitInfo h = intersect(T, r);
          if (h.bHit && h.k < bestK && h.k > 0.0001f) {
            bestK = h.k;
            kept = t;
          }
        }
        if ((int)(bvh[9 * curr]) != -1) push(&S, (int)(bvh[9 * curr]));
        if ((int)(bvh[9 * curr + 1]) != -1) push(&S, (int)(bvh[9 * curr + 1]));
      }
    }
    size_t o = (size_t)(i - i0);
    dir_out[3 * o + 0] = r.dir.x;
    dir_out[3 * o + 1] = r.dir.y;
    dir_out[3 * o + 2] = r.dir.z;
    hit_out[o] = H.bHit ? 1 : 0;
    k_out[o] = H.k;
    mat_out[o] = H.mat;
    tri_out[o] = H.bHit ? kept : -1;
    n_out[3 * o + 0] = H.n.x;
    n_out[3 * o + 1] = H.n.y;
    n_out[3 * o + 2] = H.n.z;
  }
}

// Reference `ImgProcessing` kernel over work-items [0, global) (KernelLauncher.py:101-102).
void ref_img_processing(const float *in, float *out, int N, int global) {
  for (int i = 0; i < global; ++i) {
    g_global_id = i;
    ImgProcessing((float *)in, out, N);
  }
}

// n draws of the reference's rand() for one pixel, wired exactly as the kernel
// wires it: Raytracing passes (&seed0,&seed1) to naiveGI (Raytracing.cl:205-206),
// naiveGI passes (seed1, seed0) to the samplers (:61-69), which call
// rand(seed0, seed1) with their own parameter names (MathLib.cl:316-317).
void ref_rand_stream(int pixel, int imgSize, int n, float *out) {
  unsigned int seed0 = pixel % imgSize;
  unsigned int seed1 = pixel / imgSize;
  unsigned int *k_seed0 = &seed0, *k_seed1 = &seed1;        // kernel scope
  unsigned int *s_seed0 = k_seed1, *s_seed1 = k_seed0;      // sampler scope (swapped by naiveGI)
  for (int j = 0; j < n; ++j) out[j] = rand(s_seed0, s_seed1);
}

int ref_has_counters(void) {
#ifdef CLREF_COUNTERS
  return 1;
#else
  return 0;
#endif
}
}
