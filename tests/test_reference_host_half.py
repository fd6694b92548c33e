"""The reference's own scene importer (FileManager.Scene) run UNCHANGED against this repo's drop-ins:
`compat/` stand-ins for the packages it imports but this image lacks, and the native `BVH` class in place of
BVH.py.  The buffers it produces must be the committed fixtures (which came from the reference's BVH.py).
Needs the reference checkout, so it runs in the authoring container only."""
import contextlib
import importlib
import io
import os
import shutil
import sys
import warnings

import numpy as np
import pytest

from tests import fixtures

REF = os.environ.get("B200RT_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "ensem3a_openclraytracer_b200", "compat")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ObjFiles")), reason="reference checkout absent")


@pytest.fixture()
def reference_filemanager(monkeypatch):
    """import FileManager from the reference with compat/ on the path and BVH -> the native drop-in."""
    import ensem3a_openclraytracer_b200  # noqa: F401
    native_bvh = sys.modules["ensem3a_openclraytracer_b200.BVH"]   # the module (the package attribute is the class)
    saved = {k: sys.modules.get(k) for k in ("FileManager", "BVH", "pywavefront", "matplotlib", "matplotlib.pyplot",
                                             "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d")}
    for k in saved:
        sys.modules.pop(k, None)
    monkeypatch.syspath_prepend(COMPAT)
    monkeypatch.syspath_prepend(REF)
    monkeypatch.setattr(sys, "dont_write_bytecode", True)
    sys.modules["BVH"] = native_bvh          # `from BVH import *` (FileManager.py:9) now binds the native class
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fm = importlib.import_module("FileManager")
    yield fm
    for k, v in saved.items():
        sys.modules.pop(k, None)
        if v is not None:
            sys.modules[k] = v


@pytest.mark.parametrize("obj,fixture", [("Cornell box", "cornell"), ("protoEnsem", "proto"), ("Serre_leger", "serre")])
def test_scene_importer_with_drop_ins(reference_filemanager, tmp_path, obj, fixture):
    src = os.path.join(REF, "ObjFiles", obj + ".obj")
    shutil.copyfile(src, tmp_path / (obj + ".obj"))
    shutil.copyfile(src.replace(".obj", ".ini"), tmp_path / (obj + ".ini"))
    os.chmod(tmp_path / (obj + ".ini"), 0o644)
    with contextlib.redirect_stdout(io.StringIO()):
        scene = reference_filemanager.Scene(str(tmp_path / (obj + ".obj")), True, None)
    want = fixtures.load_scene(fixture)
    assert type(scene.BVH).__module__.endswith("ensem3a_openclraytracer_b200.BVH")
    for key, got in (("V_p", scene.V_p), ("V_n", scene.V_n), ("V_uv", scene.V_uv), ("faceData", scene.faceData),
                     ("materialData", scene.materialData), ("BVH", scene.BVH.exportArray)):
        got = np.ascontiguousarray(got, dtype=want[key].dtype)
        assert got.shape == want[key].shape, key
        assert np.array_equal(got.view(np.uint32), want[key].view(np.uint32)), key
