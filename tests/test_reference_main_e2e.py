"""The reference's own driver, main.main(scene) (main.py:14-111; UI.py:100 calls it), executed UNCHANGED on the
drop-ins — stand-in pyopencl, drop-in KernelLauncher, drop-in BVH — and its output/out.png compared with the oracle's
image quantised the way FileManager.saveImg quantises (FileManager.py:334-338).  Needs the reference checkout (set
B200RT_REFERENCE_ROOT on a box that has it elsewhere) and a GPU."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle
from tests import fixtures

REF = os.environ.get("B200RT_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "main.py")), reason="reference checkout absent")]


@pytest.mark.parametrize("scene,fixture,res,spp", [("Cornell box", "cornell", 96, 6), ("protoEnsem", "proto", 64, 4)])
def test_reference_main_runs_unchanged_and_writes_the_oracle_image(tmp_path, scene, fixture, res, spp):
    from PIL import Image
    # a fresh interpreter: main.py / FileManager.py are imported under their own top-level names
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_main.py"), REF, str(tmp_path), scene,
                          str(res), str(spp)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    info = json.loads(out.stdout.strip().splitlines()[-1])
    assert info["launcher_module"] == "ensem3a_openclraytracer_b200.KernelLauncher"
    assert info["bvh_module"] == "ensem3a_openclraytracer_b200.BVH"
    png = np.asarray(Image.open(tmp_path / "output" / "out.png").convert("RGB"))
    assert png.shape == (res, res, 3)
    # the oracle on the committed fixture of the same scene (= the buffers the reference's own importer produces,
    # tests/test_reference_host_half.py), the environment main.py opened, the .ini's camera / sun / bounce settings
    sc = fixtures.load_scene(fixture)
    ibl = np.asarray(Image.open(tmp_path / "IBL" / "Arches_E_PineTree_8k.jpg").convert("RGBA"))
    cam, env = fixtures.cam_env(sc["params"], res)
    bounce = int(sc["params"]["maxBounce"])
    ref, _ = oracle.render(sc, cam, env, res * res, spp, bounce, ibl)
    want = (ref.reshape(res, res, 3) * 255).astype("uint8")          # FileManager.saveImg
    assert np.array_equal(png, want), f"{np.count_nonzero(png != want)} of {png.size} bytes differ"
