"""b200rt_build_bvh must emit the reference BVH.py's exportArray bit for bit.

The expected arrays in tests/golden/scene_*.npz were produced by the reference's own BVH.py
(tests/golden/make_golden.py runs FileManager.Scene unchanged)."""
import numpy as np
import pytest

import ensem3a_openclraytracer_b200 as rt
from tests import fixtures

SCENES = ["single", "cornell", "proto", "furnace", "serre", "monkey"]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", SCENES)
def test_builder_reproduces_bvh_py(name):
    sc = fixtures.load_scene(name)
    got, depth = rt.build_bvh(sc["faceData"], sc["V_p"], return_depth=True)
    want = np.asarray(sc["BVH"], np.float32)
    assert got.shape == want.shape
    diff = bits(got) != bits(want)
    # The one word BVH.py itself does not define: a box MINIMUM that is a zero occurring with both signs among the
    # node's vertices — np.min's result then depends on the SIMD width NumPy was dispatched to.  Everything else
    # (topology, numbering, every non-zero coordinate, every maximum) must match bit for bit.
    col = np.arange(want.size) % 9
    allowed = diff & (got == 0) & (want == 0) & (col >= 2) & (col <= 4)
    assert np.array_equal(diff, allowed), f"{(diff & ~allowed).sum()} of {want.size} words differ"
    assert allowed.sum() <= 2
    assert depth >= 1 or want.size == 9


def test_builder_reproduces_bvh_py_on_the_synthetic_mesh():
    """tests/golden/bvh_heightfield40.npz = the reference's BVH.py run on tests.synthetic.height_field(40)
    (tests/golden/make_bvh_golden.py) — the generator of BASELINE config 5 at a size BVH.py finishes."""
    import os
    from tests.synthetic import height_field
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bvh_heightfield40.npz"))
    vp, face = height_field(int(g["quads"]), seed=int(g["seed"]))
    assert np.array_equal(bits(rt.build_bvh(face, vp)), bits(g["BVH"]))


def test_drop_in_class():
    sc = fixtures.load_scene("cornell")
    b = rt.BVH(sc["faceData"], sc["V_p"])          # reference: BVH(faceData, V_p).exportArray (FileManager.py:245)
    assert b.exportArray.dtype == np.float32 and b.exportArray.ndim == 1
    assert np.array_equal(bits(b.exportArray), bits(sc["BVH"]))
    assert b.NodeCounter == b.exportArray.size // 9


def test_is_deterministic_under_threads():
    sc = fixtures.load_scene("monkey")
    a = rt.build_bvh(sc["faceData"], sc["V_p"])
    for _ in range(3):
        assert np.array_equal(bits(a), bits(rt.build_bvh(sc["faceData"], sc["V_p"])))


def test_tree_invariants_on_a_synthetic_mesh():
    """A jittered height field (the generator of BASELINE config 5, small): 2n-1 nodes, every triangle in exactly
    one leaf, children numbered pairwise, child boxes nested in their parent's."""
    from tests.synthetic import height_field
    vp, face = height_field(40, seed=0)
    bvh = rt.build_bvh(face, vp).reshape(-1, 9)
    n = face.size // 10
    assert bvh.shape[0] == 2 * n - 1
    leaves = bvh[bvh[:, 8] >= 0]
    assert sorted(leaves[:, 8].astype(int).tolist()) == list(range(n))
    inner = bvh[bvh[:, 8] < 0]
    assert np.all(inner[:, 1] == inner[:, 0] + 1)
    for row in inner:
        for ch in (int(row[0]), int(row[1])):
            assert np.all(bvh[ch, 2:5] >= row[2:5]) and np.all(bvh[ch, 5:8] <= row[5:8])


def test_degenerate_input_is_an_error_not_a_hang():
    # two identical triangles: coincident centroids, BVH.py recurses without end on this
    vp = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], np.float32)
    face = np.array([0, 0, 0, 0, 0, 0, 0, 0, 1, 2] * 2, np.int32)
    with pytest.raises(rt.B200RTError):
        rt.build_bvh(face, vp)
