"""CPU check of the round-2 traversal data: a plain-Python restatement of the fast traversal (csrc/rt_trace.cuh,
closest_hit_fast) walks the node records b200rt_set_scene would upload — the culling tree of csrc/cull_tree.cpp, planes
quantised onto the 2^15 grid — and must find, for every ray, exactly the hit of the reference's own walk over BVH.py's
array (the oracle).  What this pins without a GPU: the culling tree holds every triangle under boxes that enclose it,
the conservative slab test with its margin never loses a candidate, ties go to the reference's first visit (ranks), and
the winner's leaf box passes the exact test.  The walk here culls MORE eagerly than the kernels do (exact distances on
the stack instead of the coded ones, leaves tested on the spot instead of parked), so passing here implies the kernels'
walk visits a superset.  Test infrastructure, like oracle/: nothing in the product imports it."""
import numpy as np
import pytest

from oracle import oracle
from tests import fixtures
from tests.test_repack import probe

f32 = np.float32
EPS = f32(0.0000001)


def fma(a, b, c):
    """binary32 fused multiply-add: the product of two binary32 numbers is exact in binary64"""
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


class Walk:
    def __init__(self, sc):
        self.nodes, self.info = probe(sc, own_tree=True)
        i = self.info
        assert i["canonical"] and i["fast_ok"]
        self.nf4 = i["node_f4"]
        self.base, self.pitch = i["grid_base"].astype(f32), i["grid_pitch"].astype(f32)
        self.cmax, self.cull_abs = f32(i["cmax"]), f32(i["cull_abs"])
        vp = sc["V_p"].reshape(-1, 3).astype(f32)
        face = sc["faceData"].reshape(-1, 10)
        self.A = vp[face[:, 7]]
        self.e1 = vp[face[:, 8]] - self.A          # MathLib.cl:129-130
        self.e2 = vp[face[:, 9]] - self.A
        self.rank = i["ranks"]
        bvh = sc["BVH"].reshape(-1, 9)
        leaf = bvh[bvh[:, 8] != -1]
        self.lbox = np.zeros((len(face), 6), f32)
        self.lbox[leaf[:, 8].astype(int)] = leaf[:, 2:8]
        self.qmin = (self.nodes[:, :6] & 0xffff).astype(np.float64)
        self.qmax = (self.nodes[:, :6] >> 16).astype(np.float64)
        refs = self.nodes[:, 6:8].astype(np.int64)
        self.refs = np.where(refs >= 1 << 31, refs - (1 << 32), refs)
        self.visits = 0
        self.tests = 0

    def ray(self, o, d):
        A, Bn, Bf, neg = [], [], [], []
        for a in range(3):
            rr = f32(1.0) / d[a]
            b = (self.base[a] - o[a]) * rr
            m = (f32(3.814697265625e-06) * (self.cmax + abs(o[a]))) * abs(rr)
            A.append(self.pitch[a] * rr)
            Bn.append(b - m)
            Bf.append(b + m)
            neg.append(bool(np.signbit(d[a])))
        return A, Bn, Bf, neg

    def slab(self, Q, qmin, qmax):
        A, Bn, Bf, neg = Q
        lo, hi = f32(-np.inf), f32(np.inf)
        for a in range(3):
            fmin, fmax_ = f32(0.5 + qmin[a] / 65536.0), f32(0.5 + qmax[a] / 65536.0)
            near = fma(fmax_ if neg[a] else fmin, A[a], Bn[a])
            far = fma(fmin if neg[a] else fmax_, A[a], Bf[a])
            lo, hi = max(lo, near), min(hi, far)
        return lo, hi

    def tri_hit(self, t, o, d):
        e1, e2, A = self.e1[t], self.e2[t], self.A[t]
        h = cross(d, e2)
        a = dot(e1, h)
        if -EPS < a < EPS:
            return None
        f = f32(1.0) / a
        s = (o[0] - A[0], o[1] - A[1], o[2] - A[2])
        u = f * dot(s, h)
        if u < 0 or u > 1:
            return None
        q = cross(s, e1)
        v = f * dot(d, q)
        if v < 0 or u + v > 1:
            return None
        k = f * dot(e2, q)
        return k if k > EPS else None

    def closest_hit(self, o, d):
        """(triangle, distance, winner passed the exact leaf-box test)"""
        Q = self.ray(o, d)
        best_t, best_k, best_rank = -1, f32(1000.0), 0x7fffffff
        lim = fma(best_k, f32(1.001), self.cull_abs)
        info = self.info
        lo, hi = self.slab(Q, info["root_qmin"].astype(np.float64), info["root_qmax"].astype(np.float64))
        stack = [(0, lo)] if hi >= lo else []
        while stack:
            ref, entry = stack.pop()
            if entry > lim:
                continue
            if ref < 0:
                t = ~ref
                self.tests += 1
                k = self.tri_hit(t, o, d)
                if k is not None and k > f32(0.0001):
                    if k < best_k or (k == best_k and best_t >= 0 and self.rank[t] < best_rank):
                        best_t, best_k, best_rank = t, k, self.rank[t]
                        lim = fma(best_k, f32(1.001), self.cull_abs)
                continue
            n = ref // self.nf4
            self.visits += 1
            go = []
            for side in (0, 1):
                lo, hi = self.slab(Q, self.qmin[n, 3 * side:3 * side + 3], self.qmax[n, 3 * side:3 * side + 3])
                if hi >= lo and not lo > lim and not hi < -self.cull_abs:
                    go.append((int(self.refs[n, side]), lo))
            go.sort(key=lambda e: -e[1])     # the nearer child is popped first
            stack.extend(go)
        ok = True
        if best_t >= 0:                       # validate_hit: MathLib.cl:169-188 on the winner's leaf box
            b = self.lbox[best_t]
            tmin, tmax = f32(-np.inf), f32(np.inf)
            for a in range(3):
                t1, t2 = (b[a] - o[a]) / d[a], (b[3 + a] - o[a]) / d[a]
                tmin, tmax = max(tmin, min(t1, t2)), min(tmax, max(t1, t2))
            ok = bool(tmax >= tmin)
        return best_t, best_k, ok


def rays_for(sc, n, seed):
    """camera rays, then rays leaving their hit points in random directions (what a path tracer traces)"""
    rng = np.random.default_rng(seed)
    res = 24
    cam, _ = fixtures.cam_env(sc["params"], res)
    prim = oracle.primary(sc, cam, res * res)
    o0 = np.tile(cam[:3].astype(f32), (res * res, 1))
    rays = [np.concatenate([o0, prim["dir"].astype(f32)], 1)]
    hit = prim["tri"] >= 0
    d = prim["dir"][hit].astype(f32)
    p = o0[hit] + d / np.linalg.norm(d, axis=1, keepdims=True).astype(f32) * prim["k"][hit, None]
    while sum(len(r) for r in rays) < n:
        w = rng.normal(size=p.shape).astype(f32)
        w /= np.linalg.norm(w, axis=1, keepdims=True)
        rays.append(np.concatenate([p, w], 1).astype(f32))
    rays = np.concatenate(rays)[:n]
    return rays[(rays[:, 3:] != 0).all(1)]    # regular rays: the leaf box decides (irregular ones: validate_chain, GPU tests)


@pytest.mark.parametrize("name,n", [("cornell", 4000), ("monkey_cfg2", 3000), ("serre", 2500), ("furnace", 2500), ("proto", 2500), ("single", 1500)])
def test_walk_of_the_culling_tree_finds_the_reference_hit(name, n):
    sc = fixtures.load_scene(name)
    rays = rays_for(sc, n, seed=5)
    want_tri, want_k, cnt = oracle.trace_rays(sc, rays)
    W = Walk(sc)
    regraze = 0
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for i, r in enumerate(rays):
            o, d = tuple(r[:3]), tuple(r[3:])
            t, k, ok = W.closest_hit(o, d)
            if not ok:                      # a winner that grazes its leaf box within rounding: the kernels re-trace exactly
                regraze += 1
                continue
            assert t == want_tri[i], (name, i, t, want_tri[i], k, want_k[i])
            assert f32(k).view(np.uint32) == want_k[i].view(np.uint32), (name, i)
    assert regraze <= 2
    assert (want_tri >= 0).sum() > len(rays) // 4
    # and it is the cheaper walk: fewer box tests than the reference's order over BVH.py's tree spends
    assert 2 * W.visits < cnt["box_tests"]


@pytest.mark.parametrize("shift", [1.0e3, 1.0e5, 3.0e6])
def test_scene_far_from_the_origin(shift):
    """The Cornell box moved far away from the origin relative to its size: float32 spacing there is coarse (0.25 at 3e6),
    the grid of the node records has to widen until its end planes enclose the root box after rounding, and the margin of
    the conservative test grows with |coordinate| — the walk may visit more, it must not lose a hit."""
    import ensem3a_openclraytracer_b200 as rt
    sc = fixtures.load_scene("cornell")
    vp = sc["V_p"].reshape(-1, 3).copy()
    vp[:, 0] += f32(shift)
    vp[:, 2] -= f32(shift / 3)
    far = dict(sc, V_p=vp.reshape(-1).astype(f32))
    far["BVH"] = rt.build_bvh(far["faceData"], far["V_p"])
    far["params"] = dict(sc["params"], cam_x=str(float(sc["params"]["cam_x"]) + shift),
                         cam_z=str(float(sc["params"]["cam_z"]) - shift / 3))
    rays = rays_for(far, 1200, seed=3)
    want_tri, want_k, _ = oracle.trace_rays(far, rays)
    W = Walk(far)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        for i, r in enumerate(rays):
            t, k, ok = W.closest_hit(tuple(r[:3]), tuple(r[3:]))
            if ok:
                assert t == want_tri[i] and f32(k).view(np.uint32) == want_k[i].view(np.uint32), (shift, i)
    assert (want_tri >= 0).sum() > 300
