// BVH traversal and ray/triangle intersection for the b200rt kernels.
//
// Contract: closest_hit_*() returns, for any ray, exactly the (triangle, distance) the
// reference's rayTrace() keeps (MathLib.cl:234-288):
//   * a triangle is a candidate iff the slab test of MathLib.cl:167-190 — six TRUE divisions, a
//     line/box test with no clipping to t >= 0 — passes for its leaf box and for every ancestor box,
//     and Möller–Trumbore (MathLib.cl:117-160) reports k > 1e-7 with the reference's u/v rejections;
//   * among candidates with 1e-4 < k < 1000 the smallest k wins; equal k is resolved in favour of the
//     triangle the reference visits first (it pushes left then right, so it walks right-first).
//
// Two implementations:
//   closest_hit_reference  the reference's own visiting order on the as-is 9-float nodes, including
//                          its capped stack that silently drops pushes (stack.cl:21-26);
//   closest_hit_fast       front-to-back over a repacked 64-byte two-child node, sub-trees skipped when
//                          their entry distance exceeds the best hit (plus a safety margin) or when
//                          they lie wholly behind the origin, ties resolved by each triangle's
//                          precomputed rank in the reference's visiting order.  Every box decision it
//                          does take is the same exact slab test.
#pragma once
#include "rt_math.cuh"

namespace b200rt {

// ---- repacked scene -------------------------------------------------------------------------------
// node (64 B, 4 x float4), interior nodes only, root = 0:
//   q0 = Lmin.x Lmin.y Lmin.z Lmax.x
//   q1 = Lmax.y Lmax.z Rmin.x Rmin.y
//   q2 = Rmin.z Rmax.x Rmax.y Rmax.z
//   q3 = refL refR (int bits)  -  -        ref >= 0: interior node index, ref < 0: leaf of triangle ~ref
// triangle (48 B, 3 x float4):
//   t0 = A.x A.y A.z e1.x     t1 = e1.y e1.z e2.x e2.y     t2 = e2.z mat rank -   (mat, rank int bits)
//   with e1 = B - A, e2 = C - A rounded exactly as MathLib.cl:129-130 rounds them.
// normal (16 B): first-vertex normal of the triangle (MathLib.cl:151), w unused.
struct SceneView {
  const float4 *nodes;
  const float4 *tris;
  const float4 *normals;
  const float *mats;        // 6 floats per material
  // as-is reference buffers for closest_hit_reference
  const float *bvh9;
  int root_ref;             // ~tri when the whole tree is one leaf
  float root_box[6];        // min xyz, max xyz of node 0
  float cull_abs;           // absolute part of the culling margin (1e-3 x scene diagonal)
  int stack_cap;            // reference traversal
  int fast_div_ok;          // scene coordinates small enough for div_by()
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

// ---- slab test --------------------------------------------------------------------------------------
struct RayDiv {
  v3 o, d, r;   // origin, direction, RN(1/direction)
  bool fast;    // div_by() is valid for this ray
};

RT_DEV RayDiv make_raydiv(v3 o, v3 d, bool scene_ok) {
  RayDiv R;
  R.o = o; R.d = d;
  R.fast = scene_ok && div_safe(d.x) && div_safe(d.y) && div_safe(d.z) &&
           fabsf(o.x) <= 1.099511627776e12f && fabsf(o.y) <= 1.099511627776e12f && fabsf(o.z) <= 1.099511627776e12f;
  R.r = mk3(__frcp_rn(d.x), __frcp_rn(d.y), __frcp_rn(d.z));
  return R;
}

// MathLib.cl:169-188.  FAST selects how the six quotients are formed, never what they are.
template <bool FAST>
RT_DEV bool slab(const RayDiv &R, float mnx, float mny, float mnz, float mxx, float mxy, float mxz, float *tmin,
                 float *tmax) {
  float a, b, lo, hi;
  if (FAST) { a = div_by(mnx - R.o.x, R.d.x, R.r.x); b = div_by(mxx - R.o.x, R.d.x, R.r.x); }
  else { a = __fdiv_rn(mnx - R.o.x, R.d.x); b = __fdiv_rn(mxx - R.o.x, R.d.x); }
  lo = fminf(a, b);
  hi = fmaxf(a, b);
  if (FAST) { a = div_by(mny - R.o.y, R.d.y, R.r.y); b = div_by(mxy - R.o.y, R.d.y, R.r.y); }
  else { a = __fdiv_rn(mny - R.o.y, R.d.y); b = __fdiv_rn(mxy - R.o.y, R.d.y); }
  lo = fmaxf(lo, fminf(a, b));
  hi = fminf(hi, fmaxf(a, b));
  if (FAST) { a = div_by(mnz - R.o.z, R.d.z, R.r.z); b = div_by(mxz - R.o.z, R.d.z, R.r.z); }
  else { a = __fdiv_rn(mnz - R.o.z, R.d.z); b = __fdiv_rn(mxz - R.o.z, R.d.z); }
  lo = fmaxf(lo, fminf(a, b));
  hi = fminf(hi, fmaxf(a, b));
  *tmin = lo;
  *tmax = hi;
  return hi >= lo;
}

// ---- Möller–Trumbore, MathLib.cl:117-160 ---------------------------------------------------------------
// Returns true and k when the reference's intersect() sets bHit.
RT_DEV bool tri_hit(v3 o, v3 d, v3 A, v3 e1, v3 e2, float *k_out) {
  const float eps = 0.0000001f;
  v3 h = cross(d, e2);
  float a = dot(e1, h);
  if (a > -eps && a < eps) return false;
  float f = __fdiv_rn(1.0f, a);  // (float)(1.0 / (double)a) == RN(1/a): double rounding is innocuous for division
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

// ---- reference-order traversal ---------------------------------------------------------------------------
template <bool SMEM, bool STATS>
RT_DEV Hit closest_hit_reference(const SceneView &S, v3 o, v3 d, TraceCounters *cnt) {
  Hit best;
  best.tri = -1;
  best.k = 1000.0f;
  RayDiv R = make_raydiv(o, d, false);
  int stack[64];
  int top = -1;
  const int cap = S.stack_cap;
  stack[++top] = 0;
  while (top != -1) {
    int cur = stack[top--];
    const float *n = S.bvh9 + 9 * (size_t)cur;
    float tmin, tmax;
    if (STATS) cnt->box_tests++;
    if (!slab<false>(R, __ldg(n + 2), __ldg(n + 3), __ldg(n + 4), __ldg(n + 5), __ldg(n + 6), __ldg(n + 7), &tmin, &tmax))
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

// ---- fast traversal -------------------------------------------------------------------------------------------
// Per-lane stack in shared memory: entry e of lane l lives at stack[e * stride + l] (bank-conflict free),
// each entry = (node ref, entry distance).
struct LaneStack {
  float2 *base;   // already offset to this thread
  int stride;     // threads per block
};

template <bool SMEM, bool STATS, bool FAST>
RT_DEV Hit traverse_fast(const SceneView &S, const RayDiv &R, LaneStack st, TraceCounters *cnt) {
  Hit best;
  best.tri = -1;
  best.k = 1000.0f;
  int best_rank = 0x7fffffff;
  const v3 o = R.o, d = R.d;
  float tmin, tmax;
  if (STATS) cnt->box_tests++;
  if (!slab<FAST>(R, S.root_box[0], S.root_box[1], S.root_box[2], S.root_box[3], S.root_box[4], S.root_box[5], &tmin,
                  &tmax))
    return best;
  if (S.root_ref < 0) {
    if (STATS) cnt->tri_tests++;
    test_triangle<SMEM>(S, ~S.root_ref, o, d, best, best_rank);
    return best;
  }
  const float behind = -S.cull_abs;
  int cur = 0;
  int sp = 0;
  for (;;) {
    const float4 *p = S.nodes + 4 * (size_t)cur;
    float4 q0 = ld4<SMEM>(p), q1 = ld4<SMEM>(p + 1), q2 = ld4<SMEM>(p + 2), q3 = ld4<SMEM>(p + 3);
    int refL = __float_as_int(q3.x), refR = __float_as_int(q3.y);
    float tminL, tmaxL, tminR, tmaxR;
    if (STATS) cnt->box_tests += 2;
    bool goL = slab<FAST>(R, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tminL, &tmaxL);
    bool goR = slab<FAST>(R, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tminR, &tmaxR);
    float lim = best.k * 1.001f + S.cull_abs;
    goL = goL && !(tminL > lim) && !(tmaxL < behind);
    goR = goR && !(tminR > lim) && !(tmaxR < behind);
    // leaves first (nearer one first so the other can be culled by the new limit)
    if (goL && goR && refL < 0 && refR < 0 && tminR < tminL) {
      if (STATS) cnt->tri_tests++;
      test_triangle<SMEM>(S, ~refR, o, d, best, best_rank);
      goR = false;
      lim = best.k * 1.001f + S.cull_abs;
      goL = !(tminL > lim);
    }
    if (goL && refL < 0) {
      if (STATS) cnt->tri_tests++;
      test_triangle<SMEM>(S, ~refL, o, d, best, best_rank);
      goL = false;
      lim = best.k * 1.001f + S.cull_abs;
      goR = goR && !(tminR > lim);
    }
    if (goR && refR < 0) {
      if (STATS) cnt->tri_tests++;
      test_triangle<SMEM>(S, ~refR, o, d, best, best_rank);
      goR = false;
      lim = best.k * 1.001f + S.cull_abs;
      goL = goL && !(tminL > lim);
    }
    if (goL && goR) {
      bool leftNear = tminL <= tminR;
      int farRef = leftNear ? refR : refL;
      float farT = leftNear ? tminR : tminL;
      st.base[sp * st.stride] = make_float2(__int_as_float(farRef), farT);
      ++sp;
      cur = leftNear ? refL : refR;
    } else if (goL) {
      cur = refL;
    } else if (goR) {
      cur = refR;
    } else {
      bool found = false;
      while (sp > 0) {
        --sp;
        float2 e = st.base[sp * st.stride];
        if (!(e.y > best.k * 1.001f + S.cull_abs)) {
          cur = __float_as_int(e.x);
          found = true;
          break;
        }
      }
      if (!found) break;
    }
  }
  return best;
}

template <bool SMEM, bool STATS>
RT_DEV Hit closest_hit_fast(const SceneView &S, v3 o, v3 d, LaneStack st, TraceCounters *cnt) {
  RayDiv R = make_raydiv(o, d, S.fast_div_ok != 0);
  if (R.fast) return traverse_fast<SMEM, STATS, true>(S, R, st, cnt);
  return traverse_fast<SMEM, STATS, false>(S, R, st, cnt);
}

// ---- the same fast traversal, one node per call -------------------------------------------------------------------
// k_paths keeps every lane's traversal in registers and advances all of them one node at a time, so that
// a warp can leave the traversal loop as soon as too few of its lanes are still traversing, shade the
// finished ones into new rays and come back — instead of idling until its longest ray is done.
struct Trav {
  RayDiv R;
  Hit best;
  int best_rank;
  int cur, sp;
  bool active;
};

template <bool SMEM, bool STATS>
RT_DEV void trav_begin(const SceneView &S, Trav &T, v3 o, v3 d, TraceCounters *cnt) {
  T.R = make_raydiv(o, d, S.fast_div_ok != 0);
  T.best.tri = -1;
  T.best.k = 1000.0f;
  T.best_rank = 0x7fffffff;
  T.cur = 0;
  T.sp = 0;
  float tmin, tmax;
  if (STATS) cnt->box_tests++;
  bool hit = T.R.fast ? slab<true>(T.R, S.root_box[0], S.root_box[1], S.root_box[2], S.root_box[3], S.root_box[4],
                                   S.root_box[5], &tmin, &tmax)
                      : slab<false>(T.R, S.root_box[0], S.root_box[1], S.root_box[2], S.root_box[3], S.root_box[4],
                                    S.root_box[5], &tmin, &tmax);
  T.active = hit;
  if (hit && S.root_ref < 0) {
    if (STATS) cnt->tri_tests++;
    test_triangle<SMEM>(S, ~S.root_ref, o, d, T.best, T.best_rank);
    T.active = false;
  }
}

template <bool SMEM, bool STATS>
RT_DEV void trav_step(const SceneView &S, Trav &T, LaneStack st, TraceCounters *cnt) {
  const float4 *p = S.nodes + 4 * (size_t)T.cur;
  float4 q0 = ld4<SMEM>(p), q1 = ld4<SMEM>(p + 1), q2 = ld4<SMEM>(p + 2), q3 = ld4<SMEM>(p + 3);
  int refL = __float_as_int(q3.x), refR = __float_as_int(q3.y);
  float tminL, tmaxL, tminR, tmaxR;
  bool goL, goR;
  if (STATS) cnt->box_tests += 2;
  if (T.R.fast) {
    goL = slab<true>(T.R, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tminL, &tmaxL);
    goR = slab<true>(T.R, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tminR, &tmaxR);
  } else {
    goL = slab<false>(T.R, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tminL, &tmaxL);
    goR = slab<false>(T.R, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tminR, &tmaxR);
  }
  const float behind = -S.cull_abs;
  float lim = T.best.k * 1.001f + S.cull_abs;
  goL = goL && !(tminL > lim) && !(tmaxL < behind);
  goR = goR && !(tminR > lim) && !(tmaxR < behind);
  // leaves are tested at once (one triangle each); the nearer one first so that the farther can still be culled
  bool leafL = goL && refL < 0, leafR = goR && refR < 0;
  if (leafL || leafR) {
    bool rightFirst = leafR && (!leafL || tminR < tminL);
    int t0 = rightFirst ? ~refR : ~refL;
    if (STATS) cnt->tri_tests++;
    test_triangle<SMEM>(S, t0, T.R.o, T.R.d, T.best, T.best_rank);
    lim = T.best.k * 1.001f + S.cull_abs;
    if (leafL && leafR) {
      float tOther = rightFirst ? tminL : tminR;
      if (!(tOther > lim)) {
        if (STATS) cnt->tri_tests++;
        test_triangle<SMEM>(S, rightFirst ? ~refL : ~refR, T.R.o, T.R.d, T.best, T.best_rank);
        lim = T.best.k * 1.001f + S.cull_abs;
      }
    }
    goL = goL && !leafL && !(tminL > lim);
    goR = goR && !leafR && !(tminR > lim);
  }
  if (goL && goR) {
    bool leftNear = tminL <= tminR;
    st.base[T.sp * st.stride] = make_float2(__int_as_float(leftNear ? refR : refL), leftNear ? tminR : tminL);
    ++T.sp;
    T.cur = leftNear ? refL : refR;
  } else if (goL) {
    T.cur = refL;
  } else if (goR) {
    T.cur = refR;
  } else {
    bool found = false;
    while (T.sp > 0) {
      --T.sp;
      float2 e = st.base[T.sp * st.stride];
      if (!(e.y > T.best.k * 1.001f + S.cull_abs)) {
        T.cur = __float_as_int(e.x);
        found = true;
        break;
      }
    }
    T.active = found;
  }
}

// Leaves found by a node step are not tested on the spot (only a few lanes of a warp reach a leaf in the
// same turn) but parked in a four-entry per-lane array in shared memory (entry e of lane l at parks[e * stride]);
// the warp tests parked triangles together once enough lanes hold one.  Last in, first out; of a pair of
// leaves the farther is parked first so that the nearer is tested first.
constexpr int kParkCap = 4;

// both children through the true-division slab test; out of line: only rays with a zero, denormal or huge
// direction component come here
// (everything by value so that no caller state is forced into local memory).  Returns (tminL, tminR, goL, goR)
// with the verdicts as 1.0f / 0.0f and the "wholly behind the origin" cull already applied through `behind`.
__device__ __noinline__ float4 slab_pair_slow(v3 o, v3 d, float4 q0, float4 q1, float4 q2, float behind) {
  RayDiv R;
  R.o = o; R.d = d; R.r = mk3(0.0f, 0.0f, 0.0f); R.fast = false;
  float tminL, tmaxL, tminR, tmaxR;
  bool goL = slab<false>(R, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tminL, &tmaxL);
  bool goR = slab<false>(R, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tminR, &tmaxR);
  goL = goL && !(tmaxL < behind);
  goR = goR && !(tmaxR < behind);
  return make_float4(tminL, tminR, goL ? 1.0f : 0.0f, goR ? 1.0f : 0.0f);
}

// one node; leaf children that pass their box test are parked instead of tested (needs pn <= kParkCap - 2 on entry)
template <bool SMEM, bool STATS>
RT_DEV void trav_step_park(const SceneView &S, Trav &T, int &pn, uint32_t *parks, int pstride, LaneStack st,
                           TraceCounters *cnt) {
  const float4 *p = S.nodes + 4 * (size_t)T.cur;
  float4 q0 = ld4<SMEM>(p), q1 = ld4<SMEM>(p + 1), q2 = ld4<SMEM>(p + 2), q3 = ld4<SMEM>(p + 3);
  const int refL = __float_as_int(q3.x), refR = __float_as_int(q3.y);
  float tminL, tminR;
  bool goL, goR;
  if (STATS) cnt->box_tests += 2;
  if (T.R.fast) {
    float tmaxL, tmaxR;
    goL = slab<true>(T.R, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, &tminL, &tmaxL);
    goR = slab<true>(T.R, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, &tminR, &tmaxR);
    goL = goL && !(tmaxL < -S.cull_abs);
    goR = goR && !(tmaxR < -S.cull_abs);
  } else {
    float4 r = slab_pair_slow(T.R.o, T.R.d, q0, q1, q2, -S.cull_abs);
    tminL = r.x; tminR = r.y;
    goL = r.z != 0.0f; goR = r.w != 0.0f;
  }
  const float lim = T.best.k * 1.001f + S.cull_abs;
  goL = goL && !(tminL > lim);
  goR = goR && !(tminR > lim);
  const bool leafL = goL && refL < 0, leafR = goR && refR < 0;
  const bool both = leafL && leafR;
  // parked last = tested first: the nearer leaf goes last
  const bool rightNearer = both ? (tminR < tminL) : leafR;
  const int nearerTri = rightNearer ? ~refR : ~refL;
  const int fartherTri = rightNearer ? ~refL : ~refR;
  if (both) { parks[pn * pstride] = (uint32_t)fartherTri; ++pn; }
  if (leafL || leafR) { parks[pn * pstride] = (uint32_t)nearerTri; ++pn; }
  goL = goL && !leafL;
  goR = goR && !leafR;
  const bool leftNear = tminL <= tminR;
  if (goL && goR) {
    st.base[T.sp * st.stride] = make_float2(__int_as_float(leftNear ? refR : refL), leftNear ? tminR : tminL);
    ++T.sp;
    T.cur = leftNear ? refL : refR;
  } else if (goL || goR) {
    T.cur = goL ? refL : refR;
  } else {
    bool found = false;
    while (T.sp > 0) {
      --T.sp;
      float2 e = st.base[T.sp * st.stride];
      if (!(e.y > lim)) {
        T.cur = __float_as_int(e.x);
        found = true;
        break;
      }
    }
    T.active = found;
  }
}

// TRAV: 0 fast, 1 reference, 2 verify (both; keeps reference, counts disagreements)
template <int TRAV, bool SMEM, bool STATS>
RT_DEV Hit closest_hit(const SceneView &S, v3 o, v3 d, LaneStack st, TraceCounters *cnt, unsigned int *mismatch) {
  if (TRAV == 0) return closest_hit_fast<SMEM, STATS>(S, o, d, st, cnt);
  if (TRAV == 1) return closest_hit_reference<SMEM, STATS>(S, o, d, cnt);
  Hit a = closest_hit_reference<SMEM, STATS>(S, o, d, cnt);
  TraceCounters dummy;
  dummy.box_tests = 0; dummy.tri_tests = 0;
  Hit b = closest_hit_fast<SMEM, false>(S, o, d, st, &dummy);
  if (a.tri != b.tri || __float_as_int(a.k) != __float_as_int(b.k)) (*mismatch)++;
  return a;
}

}  // namespace b200rt
