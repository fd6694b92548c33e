"""CPU: the restatement oracle against oracle/_ref/libclref.so run LIVE (the reference's own .cl text
compiled by g++).  Complements test_oracle_golden.py with configurations the golden file does not
hold.  Skipped where oracle/_ref has not been built (it is built wherever /root/reference exists)."""
import numpy as np
import pytest

from oracle import oracle, ref_lib
from tests import fixtures

pytestmark = pytest.mark.skipif(not ref_lib.available(), reason="oracle/_ref not built")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name,res,spp,bounce", [
    ("cornell", 80, 5, 4), ("cornell", 33, 3, 0), ("monkey_cfg2", 40, 3, 4), ("serre", 56, 3, 3),
    ("furnace_cfg3", 48, 4, 4), ("proto", 50, 3, 1), ("single", 40, 2, 4),
])
def test_render_bit_exact(name, res, spp, bounce):
    sc = fixtures.load_scene(name)
    cam, env = fixtures.cam_env(sc["params"], res)
    ibl = fixtures.load_ibl("grey" if name == "furnace_cfg3" else "preview")
    want, wc = ref_lib.raytrace(sc, cam, env, res * res, spp, bounce, ibl, counters=True)
    got, gc = oracle.render(sc, cam, env, res * res, spp, bounce, ibl)
    assert np.array_equal(bits(got), bits(want))
    assert gc == wc


def test_random_cameras_bit_exact():
    rng = np.random.default_rng(11)
    sc = fixtures.load_scene("serre")
    ibl = fixtures.load_ibl()
    for _ in range(4):
        cam, env = fixtures.cam_env(sc["params"], 40)
        cam[0:3] += rng.uniform(-1, 1, 3).astype(np.float32)
        cam[3:6] = rng.uniform(-60, 60, 3).astype(np.float32)
        cam[9] = np.float32(rng.uniform(0.4, 1.6))
        env[0:3] = rng.uniform(-90, 90, 3).astype(np.float32)
        env[3:5] = rng.uniform(0, 2, 2).astype(np.float32)
        want, _ = ref_lib.raytrace(sc, cam, env, 40 * 40, 3, 3, ibl)
        got, _ = oracle.render(sc, cam, env, 40 * 40, 3, 3, ibl)
        assert np.array_equal(bits(got), bits(want))
        p_ref, p = ref_lib.primary(sc, cam, 40 * 40), oracle.primary(sc, cam, 40 * 40)
        assert np.array_equal(p["tri"], p_ref["tri"])
        assert np.array_equal(bits(p["k"]), bits(p_ref["k"]))
        assert np.array_equal(bits(p["dir"]), bits(p_ref["dir"]))


def test_pixel_subrange_matches_full_launch():
    sc = fixtures.load_scene("cornell")
    cam, env = fixtures.cam_env(sc["params"], 48)
    ibl = fixtures.load_ibl()
    full, _ = ref_lib.raytrace(sc, cam, env, 48 * 48, 3, 4, ibl)
    part, _ = oracle.render(sc, cam, env, 48 * 48, 3, 4, ibl, i0=500, i1=900)
    assert np.array_equal(bits(part[1500:2700]), bits(full[1500:2700]))
    assert np.all(part[:1500] == 0) and np.all(part[2700:] == 0)


def test_tonemap_and_rng_live():
    x = np.random.default_rng(0).uniform(-0.5, 2.0, 4096).astype(np.float32)
    assert np.array_equal(bits(oracle.img_processing(x, 4000)), bits(ref_lib.img_processing(x, 4000)))
    for px in (3, 1000, 123456):
        assert np.array_equal(bits(oracle.rand_stream(0, px, 1 << 20, 500)), bits(ref_lib.rand_stream(px, 1 << 20, 500)))
