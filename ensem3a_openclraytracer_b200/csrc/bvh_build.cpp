// b200rt_build_bvh — the reference's BVH.py, node for node, at native speed.
//
// BVH.py (reference repository) is the scene-preparation step the reference's README names as its main
// bottleneck (README.md:28): pure Python, every level re-scans and re-boxes all triangles of a node
// (8.8 s for 15 756 triangles).  The traversal kernels consume its output layout as-is, so a faster builder
// is only useful if it emits the IDENTICAL array: same splits, same node numbering, same float32 boxes.
//
// What BVH.py computes (file:line in the reference):
//   * centroid of a triangle = (a + b + c) / 3 in float64 on the float32 vertex values      BVH.py:30-40
//   * a node with more than one triangle is split on the axis of largest centroid VARIANCE
//     (np.var, population variance; first maximum wins) at the centroid MEAN (np.mean);
//     centroid[axis] < mean goes left, the rest right, original order preserved                BVH.py:73-113
//   * np.mean / np.var over a list of (3,1) arrays reduce along axis 0 of an (n,3,1) float64
//     array: one running sum per component, elements added in list order (the reduced axis is
//     the outer loop, so NumPy's pairwise summation does not apply); var = mean((x - mean)^2)
//   * box of a node = float32 min / max over the 3 vertices of its triangles, -/+ epsilon = 0  BVH.py:43-70
//     (so a maximum of -0.0 is exported as +0.0; where a minimum is a zero that occurs with both
//     signs among the vertices, the sign NumPy returns depends on the SIMD width it was dispatched to —
//     this builder follows NumPy's scalar loop (the later value wins); 2 of 71 415 words of the FurnaceHD
//     array differ from this container's NumPy in that sign; the slab test cannot tell the two zeros apart)
//   * both children are appended to the node list at split time (left first), then the left
//     subtree is built completely before the right one                                        BVH.py:107-109,147-153
//   * export: [childL, childR, min xyz, max xyz, tri] per node in list order, -1 = none,
//     leaves keep children -1, interior nodes tri -1                                          BVH.py:174-191
//
// Node numbering is a pure function of subtree sizes: when a node whose descendants start at id C splits
// into m_L and m_R triangles, its children are C and C+1, the left child's descendants start at C+2 and the
// right child's at C+2 + (2 m_L - 2).  Sub-trees are therefore independent tasks writing disjoint slices of the
// output, and the only serial part is each node's own running sums (their order is part of the contract).
//
// Degenerate input: if every centroid of a node lands on one side (coincident centroids), BVH.py silently
// drops the empty child, mis-numbers the other and recurses without end (BVH.py:169-172 with :109).  This
// builder reports B200RT_ERR_INVALID instead.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/b200rt.h"

namespace {

struct Builder {
  const float *vp;
  const int32_t *face;
  float *out;                       // (2n-1) * 9
  std::vector<double> cx, cy, cz;   // centroid per triangle
  std::vector<int32_t> idx, tmp;    // triangle ids in node order, scratch for the stable partition
  std::atomic<int> failed{0};
  std::atomic<int> max_depth{0};

  void tri_box(int32_t t, float mn[3], float mx[3]) const {
    const int32_t *f = face + 10 * (size_t)t;
    for (int v = 7; v < 10; ++v) {
      const float *p = vp + 3 * (size_t)f[v];
      for (int k = 0; k < 3; ++k) {
        if (!(mn[k] < p[k])) mn[k] = p[k];  // np.minimum's scalar loop keeps the later of two equal values (-0.0 / +0.0)
        if (p[k] > mx[k]) mx[k] = p[k];
      }
    }
  }

  void write_node(int id, int cl, int cr, const float mn[3], const float mx[3], int tri) {
    float *o = out + 9 * (size_t)id;
    o[0] = (float)cl; o[1] = (float)cr;
    // BVH.py:62-68 subtracts / adds `epsilon = 0`: x - 0 keeps x, but -0.0 + 0 is +0.0
    o[2] = mn[0]; o[3] = mn[1]; o[4] = mn[2];
    o[5] = mx[0] + 0.0f; o[6] = mx[1] + 0.0f; o[7] = mx[2] + 0.0f;
    o[8] = (float)tri;
  }

  // node `id` holds idx[lo, hi); its descendants are numbered from `first`
  void build(size_t lo, size_t hi, int id, int first, int depth) {
    for (;;) {
      if (failed.load(std::memory_order_relaxed)) return;
      const size_t n = hi - lo;
      float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
      for (size_t i = lo; i < hi; ++i) tri_box(idx[i], mn, mx);
      if (n == 1) {
        write_node(id, -1, -1, mn, mx, idx[lo]);
        int d = max_depth.load(std::memory_order_relaxed);
        while (depth > d && !max_depth.compare_exchange_weak(d, depth)) {}
        return;
      }
      // np.mean(L, axis=0): running sums in list order, then one division
      double sx = 0.0, sy = 0.0, sz = 0.0;
      for (size_t i = lo; i < hi; ++i) {
        const int32_t t = idx[i];
        sx += cx[t]; sy += cy[t]; sz += cz[t];
      }
      const double dn = (double)n;
      const double mean[3] = {sx / dn, sy / dn, sz / dn};
      // np.var(L, axis=0): mean of (x - mean)^2, same order
      double vx = 0.0, vy = 0.0, vz = 0.0;
      for (size_t i = lo; i < hi; ++i) {
        const int32_t t = idx[i];
        const double dx = cx[t] - mean[0], dy = cy[t] - mean[1], dz = cz[t] - mean[2];
        vx += dx * dx; vy += dy * dy; vz += dz * dz;
      }
      const double var[3] = {vx / dn, vy / dn, vz / dn};
      int axis = 0;  // np.argmax: first maximum
      if (var[1] > var[axis]) axis = 1;
      if (var[2] > var[axis]) axis = 2;
      const double *c = axis == 0 ? cx.data() : (axis == 1 ? cy.data() : cz.data());
      const double pivot = mean[axis];
      // stable partition: centroid < pivot goes left
      size_t nl = 0, nr = 0;
      for (size_t i = lo; i < hi; ++i) {
        const int32_t t = idx[i];
        if (c[t] < pivot) idx[lo + nl++] = t;
        else tmp[lo + nr++] = t;
      }
      if (nl == 0 || nr == 0) {
        failed.store(1);
        return;
      }
      memcpy(&idx[lo + nl], &tmp[lo], nr * sizeof(int32_t));
      const int cl = first, cr = first + 1;
      write_node(id, cl, cr, mn, mx, -1);
      const int first_l = first + 2;
      const int first_r = first + 2 + (int)(2 * nl - 2);
      const size_t mid = lo + nl;
      // the smaller side becomes a task (when it is worth one), the other continues in this thread
      const bool left_small = nl <= nr;
      const size_t s_lo = left_small ? lo : mid, s_hi = left_small ? mid : hi;
      const int s_id = left_small ? cl : cr, s_first = left_small ? first_l : first_r;
      if (s_hi - s_lo >= 4096) {
#pragma omp task firstprivate(s_lo, s_hi, s_id, s_first, depth)
        build(s_lo, s_hi, s_id, s_first, depth + 1);
      } else {
        build(s_lo, s_hi, s_id, s_first, depth + 1);
      }
      if (left_small) { lo = mid; id = cr; first = first_r; }
      else { hi = mid; id = cl; first = first_l; }
      ++depth;
    }
  }
};

}  // namespace

extern "C" int b200rt_build_bvh(const float *vertex_p, int64_t n_vertex_p, const int32_t *face_data, int64_t n_face_data,
                                float *bvh_out, int64_t n_bvh_out, int32_t *depth_out) {
  if (!vertex_p || !face_data || !bvh_out) return B200RT_ERR_INVALID;
  if (n_vertex_p <= 0 || n_vertex_p % 3 || n_face_data <= 0 || n_face_data % 10) return B200RT_ERR_INVALID;
  const int64_t n = n_face_data / 10, nv = n_vertex_p / 3;
  if (n_bvh_out != (2 * n - 1) * 9) return B200RT_ERR_INVALID;
  if (2 * n - 1 >= (1 << 24)) return B200RT_ERR_UNSUPPORTED;  // float32-encoded indices (BVH.py:165)
  for (int64_t t = 0; t < n; ++t)
    for (int v = 7; v < 10; ++v)
      if (face_data[10 * t + v] < 0 || face_data[10 * t + v] >= nv) return B200RT_ERR_INVALID;
  Builder b;
  b.vp = vertex_p;
  b.face = face_data;
  b.out = bvh_out;
  b.cx.resize((size_t)n); b.cy.resize((size_t)n); b.cz.resize((size_t)n);
  b.idx.resize((size_t)n); b.tmp.resize((size_t)n);
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < n; ++t) {
    const int32_t *f = face_data + 10 * t;
    const float *pa = vertex_p + 3 * (size_t)f[7], *pb = vertex_p + 3 * (size_t)f[8], *pc = vertex_p + 3 * (size_t)f[9];
    b.cx[t] = (((double)pa[0] + (double)pb[0]) + (double)pc[0]) / 3.0;  // BVH.py:40
    b.cy[t] = (((double)pa[1] + (double)pb[1]) + (double)pc[1]) / 3.0;
    b.cz[t] = (((double)pa[2] + (double)pb[2]) + (double)pc[2]) / 3.0;
    b.idx[t] = (int32_t)t;
  }
#pragma omp parallel
#pragma omp single
  b.build(0, (size_t)n, 0, 1, 0);
  if (b.failed.load()) return B200RT_ERR_INVALID;
  if (depth_out) *depth_out = b.max_depth.load();
  return 0;
}
