// BVH traversal and ray/triangle intersection for the b200rt kernels.
//
// Contract: every closest-hit routine here returns, for any ray, exactly the (triangle, distance) the
// reference's rayTrace() keeps (MathLib.cl:234-288):
//   * a triangle is a CANDIDATE iff the slab test of MathLib.cl:167-190 — six TRUE divisions, a line/box
//     test with no clipping to t >= 0 — passes for its leaf box and for every ancestor box, and
//     Möller–Trumbore (MathLib.cl:117-160) reports k > 1e-7 with the reference's u/v rejections;
//   * among candidates with 1e-4 < k < 1000 the smallest k wins; equal k is resolved in favour of the
//     triangle the reference visits first (it pushes left then right, so it walks right-first).
//
// Implementations:
//   closest_hit_reference  the reference's own visiting order on the as-is 9-float nodes, including its
//                          capped stack that silently drops pushes (stack.cl:21-26);
//   closest_hit_nodrop     the same walk with a stack that never drops (the fallback of the fast path, and the
//                          path of rays whose direction has a zero / denormal / huge component);
//   closest_hit_fast       the production path.  It rests on one observation about the reference's slab test:
//
//       For a ray whose direction components are all non-zero finite numbers, if a box B' is nested in a
//       box B (B.min <= B'.min <= B'.max <= B.max per axis) then slab(B') passing implies slab(B) passing.
//       Proof: RN(p - o) and RN(x / d) are monotone in p resp. x, so per axis the floating-point interval
//       [min(t1,t2), max(t1,t2)] of B contains that of B'; hence tmin(B) <= tmin(B') and tmax(B) >= tmax(B'),
//       and tmax(B') >= tmin(B') gives tmax(B) >= tmin(B).  (No NaN can arise: x / d with d != 0.)
//
//     BVH.py builds every node box as the exact float32 min/max over the node's triangles (BVH.py:43-70), so
//     boxes are nested along every root-to-leaf chain (b200rt_set_scene verifies this; a tree that violates it
//     is walked by closest_hit_reference).  A triangle is therefore a candidate iff its LEAF box passes the
//     exact slab test and Möller–Trumbore accepts it; ancestor boxes matter for culling only.  The fast path
//     walks the tree with a CONSERVATIVE slab test (FMAs only, on boxes stored as centre / half-extent and
//     widened by a proven error margin), runs
//     the exact Möller–Trumbore on the leaves it reaches, and keeps the best hit; the winner's leaf box is
//     then put through the exact slab test once (validate_hit).  If it passes — always, except for rays that
//     graze a box edge within rounding — the winner is the reference's hit: the set searched is a superset of
//     the candidates and its minimum is a candidate.  If it fails the ray is re-traced by closest_hit_nodrop.
#pragma once
#include "rt_math.cuh"

namespace b200rt {

// ---- repacked scene -------------------------------------------------------------------------------
// node (4 x float4; 64 B apart in global memory, 80 B apart when the scene is staged in shared memory so that
// lanes reading different nodes spread over all 32 banks), interior nodes only, breadth-first, root = 0:
//   q0 = Lc.x Lc.y Lc.z Lh.x        c = box centre, h = half extent, rounded so that [c - h, c + h] encloses
//   q1 = Lh.y Lh.z Rc.x Rc.y        the child's exact [min, max] (these boxes only cull; exactness lives in the
//   q2 = Rc.z Rh.x Rh.y Rh.z        leaf boxes of `tboxes` and in the as-is array `bvh9`)
//   q3 = refL refR (int bits)  -  -        ref >= 0: float4 offset of an interior node (index x node_f4),
//                                          ref < 0: leaf of triangle ~ref
// triangle (48 B, 3 x float4):
//   t0 = A.x A.y A.z e1.x     t1 = e1.y e1.z e2.x e2.y     t2 = e2.z mat rank -   (mat, rank int bits)
//   with e1 = B - A, e2 = C - A rounded exactly as MathLib.cl:129-130 rounds them.
// normal (16 B): first-vertex normal of the triangle (MathLib.cl:151), w unused.
// leaf box (32 B, 2 x float4): min.xyz -, max.xyz -   of the leaf that holds the triangle (validate_hit).
struct SceneView {
  const float4 *nodes;
  int node_f4;              // float4 per node: 4, or 5 for a scene small enough to be staged in shared memory
  const float4 *tris;
  const float4 *normals;
  const float4 *tboxes;
  const float4 *frames;     // kFrameVec float4 per triangle (rt_shade.cuh)
  const float *mats;        // 6 floats per material
  // as-is reference buffers for closest_hit_reference
  const float *bvh9;
  int root_ref;             // ~tri when the whole tree is one leaf
  float root_ch[6];         // centre xyz, half extent xyz of node 0 (enclosing, like the node records)
  float cull_abs;           // absolute part of the culling margin (1e-3 x scene diagonal)
  float cmax;               // largest |coordinate| of any box plane
  int stack_cap;            // reference traversal
  int fast_ok;              // scene coordinates in the range the conservative test is proven for
};

struct TraceCounters {
  unsigned long long box_tests, tri_tests;
};

struct Hit {
  int tri;    // -1 = miss
  float k;    // 1000 on a miss, like H.k (MathLib.cl:239)
};

template <bool SMEM>
RT_DEV float4 ld4(const float4 *p) {
  if (SMEM) return *p;
  return __ldg(p);
}

// One node = 64 contiguous bytes.  From global memory it is fetched with two 256-bit loads (LDG.E.256, sm_100+):
// lanes of a warp sit at unrelated nodes, so the L1 cost of a node visit is one wavefront per lane per load
// instruction, and two loads instead of four halve it.
template <bool SMEM>
RT_DEV void ld_node(const float4 *p, float4 &q0, float4 &q1, float4 &q2, float4 &q3) {
  if (SMEM) {
    q0 = p[0]; q1 = p[1]; q2 = p[2]; q3 = p[3];
  } else {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w), "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w)
        : "l"(p));
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(q2.x), "=f"(q2.y), "=f"(q2.z), "=f"(q2.w), "=f"(q3.x), "=f"(q3.y), "=f"(q3.z), "=f"(q3.w)
        : "l"(p + 2));
  }
}

// ---- exact slab test, MathLib.cl:169-188 --------------------------------------------------------------
RT_DEV bool slab_exact(v3 o, v3 d, float mnx, float mny, float mnz, float mxx, float mxy, float mxz, float *tmin,
                       float *tmax) {
  float a = __fdiv_rn(mnx - o.x, d.x), b = __fdiv_rn(mxx - o.x, d.x);
  float lo = fminf(a, b), hi = fmaxf(a, b);
  a = __fdiv_rn(mny - o.y, d.y); b = __fdiv_rn(mxy - o.y, d.y);
  lo = fmaxf(lo, fminf(a, b));
  hi = fminf(hi, fmaxf(a, b));
  a = __fdiv_rn(mnz - o.z, d.z); b = __fdiv_rn(mxz - o.z, d.z);
  lo = fmaxf(lo, fminf(a, b));
  hi = fminf(hi, fmaxf(a, b));
  *tmin = lo;
  *tmax = hi;
  return hi >= lo;
}

// ---- conservative slab test ----------------------------------------------------------------------------
// Boxes are held as centre c and half extent h with [c - h, c + h] enclosing the exact [min, max].  Per ray:
// r = RN(1/d) and two constants per axis, kn = RN(-o·r) - m and kf = RN(-o·r) + m, so that
//     near = fma(-h, |r|, fma(c, r, kn))  <=  min(t_exact(min), t_exact(max))
//     far  = fma(+h, |r|, fma(c, r, kf))  >=  max(t_exact(min), t_exact(max)),     t_exact(p) = RN(RN(p - o) / d).
// Only the FMA pipe is used: no per-axis min / max is needed because |r| orders the two planes; the three
// `near` values and the three `far` values are then reduced with one 3-input max / min each.
// Error budget, with P = max(|c| + h, |o|) and u = 2^-24:  t_exact differs from the real (p - o)/d by at most
// 4u·P|1/d| (two roundings of a value of magnitude <= 2P/|d|);  the FMA chain differs from the real
// (c -+ h - o)/d by at most  2u·P|r| (r)  +  u·P|r| (o·r)  +  u·(P|r| + m) (kn / kf)  +  2u·P|r| (inner FMA)
// +  2u·P|r| (outer FMA)  <  9u·P|r|.  The margin  m = 2^-19 (cmax + |o|) |r| = 32u (cmax + |o|) |r|  covers the
// sum (13u·P|r|) more than twice.  Valid while every |d| component lies in [2^-40, 2^40] and |o|, cmax <= 2^40
// (no overflow, no denormal r).
struct RayFast {
  v3 r, kn, kf;
};

RT_DEV bool comp_ok(float d) {
  float a = fabsf(d);
  return a >= 9.094947017729282e-13f /* 2^-40 */ && a <= 1.099511627776e12f /* 2^40 */;
}

RT_DEV bool ray_is_fast(const SceneView &S, v3 o, v3 d) {
  return S.fast_ok != 0 && comp_ok(d.x) && comp_ok(d.y) && comp_ok(d.z) && fabsf(o.x) <= 1.099511627776e12f &&
         fabsf(o.y) <= 1.099511627776e12f && fabsf(o.z) <= 1.099511627776e12f;
}

RT_DEV void rayfast_axis(float o, float d, float cmax, float *r, float *kn, float *kf) {
  const float rr = __frcp_rn(d);
  const float t0 = -(o * rr);
  const float m = (1.9073486328125e-06f /* 2^-19 */ * (cmax + fabsf(o))) * fabsf(rr);
  *r = rr;
  *kn = t0 - m;
  *kf = t0 + m;
}

RT_DEV RayFast make_rayfast(const SceneView &S, v3 o, v3 d) {
  RayFast Q;
  rayfast_axis(o.x, d.x, S.cmax, &Q.r.x, &Q.kn.x, &Q.kf.x);
  rayfast_axis(o.y, d.y, S.cmax, &Q.r.y, &Q.kn.y, &Q.kf.y);
  rayfast_axis(o.z, d.z, S.cmax, &Q.r.z, &Q.kn.z, &Q.kf.z);
  return Q;
}

// lo <= tmin_exact and hi >= tmax_exact of the box enclosed by (c, h); the box certainly fails when hi < lo
RT_DEV void slab_cons(const RayFast &Q, float cx, float cy, float cz, float hx, float hy, float hz, float *lo,
                      float *hi) {
  const float nx = __fmaf_rn(-hx, fabsf(Q.r.x), __fmaf_rn(cx, Q.r.x, Q.kn.x));
  const float ny = __fmaf_rn(-hy, fabsf(Q.r.y), __fmaf_rn(cy, Q.r.y, Q.kn.y));
  const float nz = __fmaf_rn(-hz, fabsf(Q.r.z), __fmaf_rn(cz, Q.r.z, Q.kn.z));
  const float fx = __fmaf_rn(hx, fabsf(Q.r.x), __fmaf_rn(cx, Q.r.x, Q.kf.x));
  const float fy = __fmaf_rn(hy, fabsf(Q.r.y), __fmaf_rn(cy, Q.r.y, Q.kf.y));
  const float fz = __fmaf_rn(hz, fabsf(Q.r.z), __fmaf_rn(cz, Q.r.z, Q.kf.z));
  *lo = fmaxf(fmaxf(nx, ny), nz);
  *hi = fminf(fminf(fx, fy), fz);
}

// ---- Möller–Trumbore, MathLib.cl:117-160 ---------------------------------------------------------------
// Returns true and k when the reference's intersect() sets bHit.
RT_DEV bool tri_hit(v3 o, v3 d, v3 A, v3 e1, v3 e2, float *k_out) {
  const float eps = 0.0000001f;
  v3 h = cross(d, e2);
  float a = dot(e1, h);
  if (a > -eps && a < eps) return false;
  float f = __frcp_rn(a);  // (float)(1.0 / (double)a) == RN(1/a): the double rounding is innocuous for a reciprocal
  v3 s = o - A;
  float u = f * dot(s, h);
  if (u < 0.0f || u > 1.0f) return false;
  v3 q = cross(s, e1);
  float v = f * dot(d, q);
  if (v < 0.0f || u + v > 1.0f) return false;
  float k = f * dot(e2, q);
  if (k > eps) { *k_out = k; return true; }
  return false;
}

template <bool SMEM>
RT_DEV void test_triangle(const SceneView &S, int t, v3 o, v3 d, Hit &best, int &best_rank) {
  const float4 *p = S.tris + 3 * (size_t)t;
  float4 t0 = ld4<SMEM>(p), t1 = ld4<SMEM>(p + 1), t2 = ld4<SMEM>(p + 2);
  float k;
  if (tri_hit(o, d, mk3(t0.x, t0.y, t0.z), mk3(t0.w, t1.x, t1.y), mk3(t1.z, t1.w, t2.x), &k) && k > 0.0001f) {
    int rank = __float_as_int(t2.z);
    if (k < best.k || (k == best.k && best.tri >= 0 && rank < best_rank)) {
      best.k = k;
      best.tri = t;
      best_rank = rank;
    }
  }
}

// the exact slab test of the leaf box that holds triangle `tri` (global memory: once per ray)
RT_DEV bool validate_hit(const SceneView &S, v3 o, v3 d, int tri) {
  const float4 mn = __ldg(S.tboxes + 2 * (size_t)tri), mx = __ldg(S.tboxes + 2 * (size_t)tri + 1);
  float lo, hi;
  return slab_exact(o, d, mn.x, mn.y, mn.z, mx.x, mx.y, mx.z, &lo, &hi);
}

// ---- reference-order traversal ---------------------------------------------------------------------------
constexpr int kRefStack = 64;   // b200rt_set_scene routes trees that need more to REFERENCE mode with the caller's cap

template <bool SMEM, bool STATS>
RT_DEV Hit closest_hit_reference_cap(const SceneView &S, v3 o, v3 d, TraceCounters *cnt, const int cap) {
  Hit best;
  best.tri = -1;
  best.k = 1000.0f;
  int stack[kRefStack];
  int top = -1;
  stack[++top] = 0;
  while (top != -1) {
    int cur = stack[top--];
    const float *n = S.bvh9 + 9 * (size_t)cur;
    float tmin, tmax;
    if (STATS) cnt->box_tests++;
    if (!slab_exact(o, d, __ldg(n + 2), __ldg(n + 3), __ldg(n + 4), __ldg(n + 5), __ldg(n + 6), __ldg(n + 7), &tmin, &tmax))
      continue;
    int t = (int)__ldg(n + 8);
    if (t != -1) {
      if (STATS) cnt->tri_tests++;
      const float4 *p = S.tris + 3 * (size_t)t;
      float4 t0 = ld4<SMEM>(p), t1 = ld4<SMEM>(p + 1), t2 = ld4<SMEM>(p + 2);
      float k;
      if (tri_hit(o, d, mk3(t0.x, t0.y, t0.z), mk3(t0.w, t1.x, t1.y), mk3(t1.z, t1.w, t2.x), &k) && k < best.k &&
          k > 0.0001f) {
        best.k = k;
        best.tri = t;
      }
    }
    int l = (int)__ldg(n), r = (int)__ldg(n + 1);
    if (l != -1 && top != cap - 1) stack[++top] = l;
    if (r != -1 && top != cap - 1) stack[++top] = r;
  }
  return best;
}

template <bool SMEM, bool STATS>
RT_DEV Hit closest_hit_reference(const SceneView &S, v3 o, v3 d, TraceCounters *cnt) {
  return closest_hit_reference_cap<SMEM, STATS>(S, o, d, cnt, S.stack_cap);
}

// The reference's walk with a stack that cannot drop: what the fast path computes, obtained the slow way.  Out of
// line: reached only by rays with a zero / denormal / huge direction component and by the (very rare) rays whose
// fast-path winner fails validate_hit.
template <bool SMEM>
__device__ __noinline__ Hit closest_hit_nodrop(const SceneView &S, v3 o, v3 d) {
  TraceCounters dummy;
  return closest_hit_reference_cap<SMEM, false>(S, o, d, &dummy, kRefStack);
}

// ---- fast traversal ---------------------------------------------------------------------------------------------
// Per-lane stack in shared memory: entry e of lane l lives at stack[e * stride + l] (bank-conflict free),
// each entry = (node ref, conservative entry distance).
struct LaneStack {
  float2 *base;   // already offset to this thread
  int stride;     // threads per block
};

// One traversal in flight.  Kept in registers; advanced one node at a time so that a warp can interleave
// node steps, leaf tests and ray refills of its 32 lanes (k_trace), or simply looped (closest_hit_fast).
struct Trav {
  v3 o, d;
  RayFast Q;
  Hit best;
  int best_rank;
  float lim;     // sub-trees whose conservative entry distance exceeds this cannot hold a closer hit
  int cur, sp;   // cur: float4 offset of the node to visit next
  bool active;
};

RT_DEV float cull_limit(const SceneView &S, float best_k) { return best_k * 1.001f + S.cull_abs; }

// Leaves whose box passes the conservative test are not tested on the spot (only a few lanes of a warp reach a
// leaf in the same turn) but parked, at most kParkCap per lane (entry e of lane l at parks[e * stride]); the
// warp tests parked triangles together.  Last in, first out; of a pair of leaves the farther is parked first.
constexpr int kParkCap = 4;

// starts a traversal of a ray for which ray_is_fast() holds
template <bool SMEM, bool STATS>
RT_DEV void trav_begin(const SceneView &S, Trav &T, v3 o, v3 d, int &pn, uint32_t *parks, TraceCounters *cnt) {
  T.o = o; T.d = d;
  T.Q = make_rayfast(S, o, d);
  T.best.tri = -1;
  T.best.k = 1000.0f;
  T.best_rank = 0x7fffffff;
  T.lim = cull_limit(S, T.best.k);
  T.cur = 0;
  T.sp = 0;
  float lo, hi;
  if (STATS) cnt->box_tests++;
  slab_cons(T.Q, S.root_ch[0], S.root_ch[1], S.root_ch[2], S.root_ch[3], S.root_ch[4], S.root_ch[5], &lo, &hi);
  T.active = hi >= lo;
  if (T.active && S.root_ref < 0) {
    parks[0] = (uint32_t)S.root_ref;
    pn = 1;
    T.active = false;
  }
}

// one node: both child boxes through the conservative test; leaf children are parked as their (negative) refs
// (needs pn <= kParkCap - 2)
template <bool SMEM, bool STATS>
RT_DEV void trav_step(const SceneView &S, Trav &T, int &pn, uint32_t *parks, int pstride, LaneStack st,
                      TraceCounters *cnt) {
  float4 q0, q1, q2, q3;
  ld_node<SMEM>(S.nodes + T.cur, q0, q1, q2, q3);
  const int refL = __float_as_int(q3.x), refR = __float_as_int(q3.y);
  float loL, hiL, loR, hiR;
  if (STATS) cnt->box_tests += 2;
  slab_cons(T.Q, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &loL, &hiL);
  slab_cons(T.Q, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &loR, &hiR);
  const float lim = T.lim;
  const float behind = -S.cull_abs;
  const bool goL = hiL >= loL && !(loL > lim) && !(hiL < behind);
  const bool goR = hiR >= loR && !(loR > lim) && !(hiR < behind);
  const bool leafL = goL && refL < 0, leafR = goR && refR < 0;
  const bool inL = goL && refL >= 0, inR = goR && refR >= 0;
  // one comparison orders both the leaves (nearer leaf parked last = tested first) and the interior children
  const bool rNear = loR < loL;
  const int refNear = rNear ? refR : refL, refFar = rNear ? refL : refR;
  const bool both = leafL && leafR;
  if (both) { parks[pn * pstride] = (uint32_t)refFar; ++pn; }
  if (leafL || leafR) { parks[pn * pstride] = (uint32_t)(both ? refNear : (leafR ? refR : refL)); ++pn; }
  if (inL && inR) {
    st.base[T.sp * st.stride] = make_float2(__int_as_float(refFar), rNear ? loL : loR);
    ++T.sp;
    T.cur = refNear;
  } else if (inL || inR) {
    T.cur = inL ? refL : refR;
  } else {
    bool found = false;
    while (T.sp > 0) {
      --T.sp;
      const float2 e = st.base[T.sp * st.stride];
      if (!(e.y > lim)) {
        T.cur = __float_as_int(e.x);
        found = true;
        break;
      }
    }
    T.active = found;
  }
}

// a parked leaf: exact Möller–Trumbore, then the culling limit follows the best hit
template <bool SMEM>
RT_DEV void test_parked(const SceneView &S, Trav &T, uint32_t ref, v3 o, v3 d) {
  test_triangle<SMEM>(S, ~(int)ref, o, d, T.best, T.best_rank);
  T.lim = cull_limit(S, T.best.k);
}

// the whole walk by one thread (k_primary, k_trace_rays, verify mode); `parks` = kParkCap words of this thread
template <bool SMEM, bool STATS>
RT_DEV Hit closest_hit_fast(const SceneView &S, v3 o, v3 d, LaneStack st, uint32_t *parks, int pstride,
                            TraceCounters *cnt) {
  if (!ray_is_fast(S, o, d)) return closest_hit_nodrop<SMEM>(S, o, d);
  Trav T;
  int pn = 0;
  trav_begin<SMEM, STATS>(S, T, o, d, pn, parks, cnt);
  while (T.active || pn > 0) {
    if (pn > 0 && (!T.active || pn > kParkCap - 2)) {
      --pn;
      if (STATS) cnt->tri_tests++;
      test_parked<SMEM>(S, T, parks[pn * pstride], o, d);
    } else {
      trav_step<SMEM, STATS>(S, T, pn, parks, pstride, st, cnt);
    }
  }
  if (T.best.tri >= 0 && !validate_hit(S, o, d, T.best.tri)) return closest_hit_nodrop<SMEM>(S, o, d);
  return T.best;
}

// TRAV: 0 fast, 1 reference, 2 verify (both; keeps reference, counts disagreements)
template <int TRAV, bool SMEM, bool STATS>
RT_DEV Hit closest_hit(const SceneView &S, v3 o, v3 d, LaneStack st, uint32_t *parks, int pstride, TraceCounters *cnt,
                       unsigned int *mismatch) {
  if (TRAV == 0) return closest_hit_fast<SMEM, STATS>(S, o, d, st, parks, pstride, cnt);
  if (TRAV == 1) return closest_hit_reference<SMEM, STATS>(S, o, d, cnt);
  Hit a = closest_hit_reference<SMEM, STATS>(S, o, d, cnt);
  TraceCounters dummy;
  dummy.box_tests = 0; dummy.tri_tests = 0;
  Hit b = closest_hit_fast<SMEM, false>(S, o, d, st, parks, pstride, &dummy);
  if (a.tri != b.tri || __float_as_int(a.k) != __float_as_int(b.k)) (*mismatch)++;
  return a;
}

}  // namespace b200rt
