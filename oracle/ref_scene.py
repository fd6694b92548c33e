"""TEST INFRASTRUCTURE — runs the reference's own host code (FileManager.Scene + BVH.py)
unchanged to obtain the input buffers of the hot path.

Only usable where /root/reference exists (the authoring container).  The GPU box gets
the buffers as committed fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py).

`pywavefront` and `matplotlib` are not installed; FileManager.py imports both
(FileManager.py:6,10-13) but only uses, from pywavefront, the raw `v`/`vn`/`vt` records in
file order (FileManager.py:260,297-304).  The stubs below provide exactly that.
"""
import contextlib
import io
import os
import shutil
import sys
import tempfile
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("B200RT_REFERENCE_ROOT", "/root/reference")


class _Parser:
    def __init__(self):
        self.normals = []
        self.tex_coords = []


class _Wavefront:
    """File-order `v`, `vn`, `vt` records — what pywavefront exposes and FileManager reads."""

    def __init__(self, path, collect_faces=False, create_materials=True, **_):
        self.vertices = []
        self.parser = _Parser()
        self.materials = {}
        with open(path) as fh:
            for line in fh:
                tok = line.split()
                if not tok:
                    continue
                if tok[0] == "v":
                    self.vertices.append(tuple(float(t) for t in tok[1:4]))
                elif tok[0] == "vn":
                    self.parser.normals.append(tuple(float(t) for t in tok[1:4]))
                elif tok[0] == "vt":
                    self.parser.tex_coords.append(tuple(float(t) for t in tok[1:3]))


def _install_stubs():
    if "pywavefront" not in sys.modules:
        m = types.ModuleType("pywavefront")
        m.Wavefront = _Wavefront
        sys.modules["pywavefront"] = m
    for name in ("matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d",
                 "mpl_toolkits.mplot3d.art3d"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.modules["mpl_toolkits.mplot3d.art3d"].Poly3DCollection = object
    sys.modules["mpl_toolkits.mplot3d.art3d"].Line3DCollection = object


def _import_filemanager():
    _install_stubs()
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import FileManager  # noqa: the reference's module, unmodified
    return FileManager


def load_reference_scene(name, ini_overrides=None, workdir=None):
    """Scene `name` (e.g. "Cornell box") through the reference's Scene(path, True, None).

    The .obj and its .ini are copied to a scratch directory first (the reference tree is
    read-only and configReader would try to create a missing .ini next to the .obj,
    FileManager.py:356-383).  `ini_overrides` = {key: value} applied to the copied .ini by
    plain `key=value` line replacement before loading (used for the pinned config-2/3
    material variants of SURVEY.md §8d).  Returns a dict of numpy arrays + parameters.
    """
    FileManager = _import_filemanager()
    tmp = workdir or tempfile.mkdtemp(prefix="b200rt_scene_")
    src_obj = os.path.join(REFERENCE_ROOT, "ObjFiles", name + ".obj")
    obj = os.path.join(tmp, name + ".obj")
    shutil.copyfile(src_obj, obj)
    src_ini = src_obj.replace(".obj", ".ini")
    ini = obj.replace(".obj", ".ini")
    if os.path.exists(src_ini):
        shutil.copyfile(src_ini, ini)
        os.chmod(ini, 0o644)
    buf = io.StringIO()
    if ini_overrides and os.path.exists(ini):
        lines = open(ini).read().splitlines()
        seen = set()
        for i, ln in enumerate(lines):
            k = ln.split("=")[0]
            if k in ini_overrides:
                lines[i] = f"{k}={ini_overrides[k]}"
                seen.add(k)
        for k, v in ini_overrides.items():
            if k not in seen:
                lines.append(f"{k}={v}")
        open(ini, "w").write("\n".join(lines) + "\n")
    with contextlib.redirect_stdout(buf):
        scene = FileManager.Scene(obj, True, None)
        params = scene.loadParameters()
    out = {
        "V_p": np.ascontiguousarray(scene.V_p, dtype=np.float32),
        "V_n": np.ascontiguousarray(scene.V_n, dtype=np.float32),
        "V_uv": np.ascontiguousarray(scene.V_uv, dtype=np.float32),
        "faceData": np.ascontiguousarray(scene.faceData, dtype=np.int32),
        "materialData": np.ascontiguousarray(scene.materialData, dtype=np.float32),
        "lightData": np.ascontiguousarray(scene.lightData, dtype=np.int32),
        "BVH": np.ascontiguousarray(scene.BVH.exportArray, dtype=np.float32),
        "params": dict(params),
    }
    if workdir is None:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


def cam_env_from_params(params, resolution=None):
    """cam[10] / envData[5] exactly as main.py:59-61,72-73 marshals them."""
    res = int(params["resolution"]) if resolution is None else int(resolution)
    cam = np.array([float(params["cam_x"]), float(params["cam_y"]), float(params["cam_z"]),
                    float(params["cam_rx"]), float(params["cam_ry"]), float(params["cam_rz"]),
                    res, res, 1, float(params["cam_DOF"]) * (3.14 / 180)]).astype(np.float32)
    env = np.array([float(params["sun_rx"]), float(params["sun_ry"]), float(params["sun_rz"]),
                    float(params["sun_Power"]), float(params["IBL_Power"])]).astype(np.float32)
    return cam, env
