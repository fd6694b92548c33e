"""Host half of b200rt_set_scene (csrc/scene_repack.cpp) through its GPU-free probe: the 32-byte node records must
decode to boxes that ENCLOSE the exact float32 boxes of BVH.py's array (they only cull — but a box that is too small
would silently drop geometry), the tree facts (depth, stack need, visiting ranks) must not depend on which of the two
code paths derived them, and malformed input must be rejected wherever it sits in the arrays."""
import numpy as np
import pytest

from ensem3a_openclraytracer_b200 import _capi
from ensem3a_openclraytracer_b200._capi import B200RTError
from tests import fixtures


def probe(sc, bvh=None, own_tree=False):
    """own_tree=False: the node records keep the caller's topology (B200RT_CULL_TREE=0), which is what the checks
    against BVH.py's array below compare; own_tree=True: the culling tree of csrc/cull_tree.cpp (the default)."""
    import os
    old = os.environ.get("B200RT_CULL_TREE")
    os.environ["B200RT_CULL_TREE"] = "1" if own_tree else "0"
    try:
        return _capi.repack_probe(sc["V_p"], sc["V_n"], sc["faceData"], max(1, len(sc["materialData"]) // 6),
                                  sc["BVH"] if bvh is None else bvh)
    finally:
        if old is None:
            del os.environ["B200RT_CULL_TREE"]
        else:
            os.environ["B200RT_CULL_TREE"] = old


def decode(nodes, info):
    """(min plane, max plane) float64 arrays [n_inner, 2 children, 3 axes] from the packed records (rt_trace.cuh):
    one word per child and axis, grid index of the max plane << 16 | grid index of the min plane."""
    w = nodes[:, :6].astype(np.uint64).reshape(-1, 2, 3)
    qmax, qmin = w >> 16, w & 0xffff
    assert qmax.max() <= 32767 and (qmin <= qmax).all()
    base, pitch = info["grid_base"].astype(np.float64), info["grid_pitch"].astype(np.float64)
    plane = lambda q: base + (0.5 + q.astype(np.float64) / 65536.0) * pitch
    return plane(qmin), plane(qmax)


def walk_pairs(sc_bvh, nodes, info):
    """yields (record index, bvh9 node) for every interior node by following the refs from the root"""
    bvh = sc_bvh.reshape(-1, 9)
    nf4 = info["node_f4"]
    refs = nodes[:, 6:8].astype(np.int64)
    refs = np.where(refs >= 1 << 31, refs - (1 << 32), refs)
    out, stack = [], [(0, 0)]
    while stack:
        rec, node = stack.pop()
        out.append((rec, node))
        for side in (0, 1):
            ch = int(bvh[node, side])
            ref = int(refs[rec, side])
            if bvh[ch, 8] != -1:
                assert ref == ~int(bvh[ch, 8])
            else:
                assert ref >= 0 and ref % nf4 == 0
                stack.append((ref // nf4, ch))
    return out


@pytest.mark.parametrize("name", ["cornell", "monkey", "furnace", "serre", "proto", "single", "height_field"])
def test_quantised_boxes_enclose_the_exact_ones(name):
    if name == "height_field":
        import ensem3a_openclraytracer_b200 as rt
        from tests.synthetic import height_field_scene
        sc = height_field_scene(64, seed=3)
        sc["BVH"] = rt.build_bvh(sc["faceData"], sc["V_p"])
    else:
        sc = fixtures.load_scene(name)
    nodes, info = probe(sc)
    assert info["canonical"] and info["fast_ok"]
    bvh = sc["BVH"].reshape(-1, 9).astype(np.float64)
    n_tris = sc["faceData"].size // 10
    if n_tris == 1:
        assert info["n_inner"] == 0
        return
    assert info["n_inner"] == n_tris - 1
    pmin, pmax = decode(nodes, info)
    pairs = walk_pairs(sc["BVH"], nodes, info)
    assert len(pairs) == info["n_inner"]
    rec = np.array([p[0] for p in pairs])
    node = np.array([p[1] for p in pairs])
    worst = 0.0
    for side in (0, 1):
        ch = bvh[node, side].astype(int)
        mn, mx = bvh[ch, 2:5], bvh[ch, 5:8]
        lo, hi = pmin[rec, side], pmax[rec, side]
        tol = 1e-12 * (np.abs(mn) + np.abs(mx) + 1.0)          # binary64 evaluation of the test itself
        assert (lo <= mn + tol).all() and (hi >= mx - tol).all()
        worst = max(worst, float((((hi - lo) - (mx - mn)) / info["grid_pitch"]).max()))
    assert worst < 5.0 / 65536.0                                # never looser than a few grid cells
    # the root box in record units encloses node 0
    base, pitch = info["grid_base"].astype(np.float64), info["grid_pitch"].astype(np.float64)
    rlo = base + (0.5 + info["root_qmin"].astype(np.float64) / 65536.0) * pitch
    rhi = base + (0.5 + info["root_qmax"].astype(np.float64) / 65536.0) * pitch
    assert (rlo <= bvh[0, 2:5]).all() and (rhi >= bvh[0, 5:8]).all()
    ext = bvh[0, 5:8] - bvh[0, 2:5]
    assert ((rhi - rlo) <= np.maximum(ext, ext.max() / 256.0) * 1.001 + 4 * pitch / 65536.0).all()   # and not much more
    assert info["cmax"] >= np.abs(bvh[:, 2:8]).max()


@pytest.mark.parametrize("name", ["cornell", "monkey", "furnace", "serre", "proto", "height_field"])
def test_own_culling_tree_holds_every_triangle_once_inside_enclosing_boxes(name):
    """csrc/cull_tree.cpp: a different topology over the SAME leaf boxes.  Every triangle must hang under exactly one
    leaf ref, every child box must enclose the leaf boxes of BVH.py's array beneath it, and the reported depth must
    bound the levels (it sizes the traversal stacks)."""
    if name == "height_field":
        import ensem3a_openclraytracer_b200 as rt
        from tests.synthetic import height_field_scene
        sc = height_field_scene(64, seed=3)
        sc["BVH"] = rt.build_bvh(sc["faceData"], sc["V_p"])
    else:
        sc = fixtures.load_scene(name)
    nodes, info = probe(sc, own_tree=True)
    assert info["canonical"] and info["fast_ok"]
    bvh = sc["BVH"].reshape(-1, 9).astype(np.float64)
    n_tris = sc["faceData"].size // 10
    assert info["n_inner"] == n_tris - 1
    leaf = bvh[bvh[:, 8] != -1]
    tmin = np.zeros((n_tris, 3)); tmax = np.zeros((n_tris, 3))
    tmin[leaf[:, 8].astype(int)] = leaf[:, 2:5]
    tmax[leaf[:, 8].astype(int)] = leaf[:, 5:8]
    pmin, pmax = decode(nodes, info)
    nf4 = info["node_f4"]
    refs = nodes[:, 6:8].astype(np.int64)
    refs = np.where(refs >= 1 << 31, refs - (1 << 32), refs)
    seen = np.zeros(n_tris, int)
    # post-order: exact box of every sub-tree = union of the leaf boxes beneath
    sub = {}
    deepest = 0
    stack = [(0, 0, False)]
    while stack:
        rec, level, done = stack.pop()
        if not done:
            stack.append((rec, level, True))
            for side in (0, 1):
                r = int(refs[rec, side])
                if r >= 0:
                    assert r % nf4 == 0 and 0 < r // nf4 < info["n_inner"]
                    stack.append((r // nf4, level + 1, False))
            continue
        lo = np.full(3, np.inf); hi = np.full(3, -np.inf)
        for side in (0, 1):
            r = int(refs[rec, side])
            if r < 0:
                t = ~r
                seen[t] += 1
                clo, chi = tmin[t], tmax[t]
                deepest = max(deepest, level + 1)
            else:
                clo, chi = sub.pop(r // nf4)
            assert (pmin[rec, side] <= clo).all() and (pmax[rec, side] >= chi).all()
            assert ((pmax[rec, side] - pmin[rec, side]) - (chi - clo) <= 5.0 / 65536.0 * info["grid_pitch"]).all()
            lo = np.minimum(lo, clo); hi = np.maximum(hi, chi)
        sub[rec] = (lo, hi)
    assert (seen == 1).all()
    assert deepest == info["cull_depth"]


def coincident_scene(n):
    """n copies of one triangle under a balanced hand-made tree: every leaf box and every centroid coincide (BVH.py itself
    never terminates on such input, BVH.py:102-109 — but a caller may hand over any valid array)."""
    vp = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0.5], np.float32)
    vn = np.array([0, 0, 1], np.float32)
    face = np.tile(np.array([0, 0, 0, 0, 0, 0, 0, 0, 1, 2], np.int32), n)
    box = [0.0, 0.0, 0.0, 1.0, 1.0, 0.5]
    nodes = []

    def build(tris):
        me = len(nodes)
        nodes.append(None)
        if len(tris) == 1:
            nodes[me] = [-1, -1] + box + [tris[0]]
        else:
            l = build(tris[:len(tris) // 2])
            r = build(tris[len(tris) // 2:])
            nodes[me] = [l, r] + box + [-1]
        return me

    build(list(range(n)))
    return dict(V_p=vp, V_n=vn, faceData=face, materialData=np.array([1, 1, 1, 1, 0.5, 1], np.float32),
                BVH=np.array(nodes, np.float32).ravel())


def test_own_culling_tree_on_coincident_triangles():
    n = 1000
    sc = coincident_scene(n)
    nodes, info = probe(sc, own_tree=True)
    assert info["canonical"] and info["n_inner"] == n - 1
    refs = nodes[:, 6:8].astype(np.int64)
    refs = np.where(refs >= 1 << 31, refs - (1 << 32), refs)
    leaves = np.sort(~refs[refs < 0])
    assert np.array_equal(leaves, np.arange(n))
    assert info["cull_depth"] <= 12          # nothing to split by: halved by index, depth = ceil(log2 n) + 1 at most
    pmin, pmax = decode(nodes, info)
    assert (pmin <= np.array([0, 0, 0])).all() and (pmax >= np.array([1, 1, 0.5])).all()


def renumber(bvh9):
    """the same tree with children numbered BELOW their parents (root stays 0): forces the general serial walk"""
    b = bvh9.reshape(-1, 9)
    n = b.shape[0]
    new_id = np.arange(n)
    new_id[1:] = n - np.arange(1, n)
    out = np.empty_like(b)
    for old in range(n):
        r = b[old].copy()
        for k in (0, 1):
            if r[k] != -1:
                r[k] = new_id[int(r[k])]
        out[new_id[old]] = r
    return out.reshape(-1)


@pytest.mark.parametrize("name", ["cornell", "proto", "furnace"])
def test_id_ordered_fast_path_equals_the_general_walk(name):
    sc = fixtures.load_scene(name)
    nodes_a, a = probe(sc)
    nodes_b, b = probe(sc, renumber(sc["BVH"]))
    for key in ("n_inner", "node_f4", "depth", "ref_stack_need", "canonical", "fast_ok", "cmax", "cull_abs"):
        assert a[key] == b[key], key
    assert np.array_equal(a["ranks"], b["ranks"])
    assert np.array_equal(nodes_a, nodes_b)              # same pre-order over sibling pairs, same quantisation
    assert sorted(a["ranks"]) == list(range(sc["faceData"].size // 10))


def test_tree_facts_match_the_survey():
    """max tree depth / max stack of the reference walk for the shipped scenes."""
    for name, depth, need in (("cornell", 7, 5), ("monkey", 18, 14), ("furnace", 13, 13), ("serre", 18, 14)):
        _, info = probe(fixtures.load_scene(name))
        assert (info["depth"], info["ref_stack_need"]) == (depth, need), name


def test_malformed_input_is_rejected_wherever_it_sits():
    sc = fixtures.load_scene("proto")
    n_nodes = sc["BVH"].size // 9
    for key, bad in (("V_p", np.nan), ("V_p", np.inf), ("BVH", np.nan), ("BVH", -np.inf)):
        arr = sc[key].copy()
        if key == "V_p":
            arr[3 * int(sc["faceData"][10 * (sc["faceData"].size // 20) + 8]) + 1] = bad     # a vertex some triangle uses
        else:
            arr[9 * (n_nodes // 2) + 4] = bad
        broken = dict(sc)
        broken[key] = arr
        with pytest.raises(B200RTError):
            probe(broken)
    for mutate in ("cycle", "shared_child", "child_out_of_range", "tri_out_of_range", "face_index"):
        broken = dict(sc)
        bvh = sc["BVH"].copy()
        if mutate == "cycle":
            bvh[9 * 5] = 0.0                                     # node 5's left child is the root
        elif mutate == "shared_child":
            bvh[9 * 0 + 1] = bvh[9 * 0]                          # both children of the root are the same node
        elif mutate == "child_out_of_range":
            bvh[9 * 3 + 1] = float(n_nodes + 7)
        elif mutate == "tri_out_of_range":
            leaf = int(np.nonzero(bvh.reshape(-1, 9)[:, 8] != -1)[0][3])
            bvh[9 * leaf + 8] = float(sc["faceData"].size)
        else:
            face = sc["faceData"].copy()
            face[10 * 11 + 9] = -4
            broken["faceData"] = face
        broken["BVH"] = bvh
        with pytest.raises(B200RTError):
            probe(broken)


def test_non_canonical_trees_are_flagged_not_rejected():
    """A node that keeps a triangle AND children, or a child box poking out of its parent: legal for the reference's walk,
    so accepted — but only the reference-order traversal may be used (canonical = False)."""
    sc = fixtures.load_scene("proto")
    bvh = sc["BVH"].copy().reshape(-1, 9)
    inner = int(np.nonzero(bvh[:, 8] == -1)[0][4])
    child = int(bvh[inner, 0])
    bvh[child, 5] += 1.0                                         # child's max.x beyond the parent's
    _, info = probe(sc, bvh.reshape(-1))
    assert not info["canonical"]
