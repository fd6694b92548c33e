"""Committed fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SCENE_NAMES = ["cornell", "monkey", "furnace", "serre", "proto", "single"]
_cache = {}


def load_scene(name):
    """name: cornell | monkey | monkey_cfg2 | furnace | furnace_cfg3 | serre | proto | single"""
    if name in _cache:
        return _cache[name]
    base = name.split("_cfg")[0]
    z = np.load(os.path.join(GOLDEN, f"scene_{base}.npz"))
    sc = {k: np.ascontiguousarray(z[k]) for k in ("V_p", "V_n", "V_uv", "faceData", "materialData", "lightData", "BVH")}
    sc["params"] = json.loads(bytes(z["params_json"]).decode())
    if name == "monkey_cfg2":
        m = np.load(os.path.join(GOLDEN, "scene_monkey_cfg2_materials.npz"))
        sc["materialData"] = np.ascontiguousarray(m["materialData"])
        sc["lightData"] = np.ascontiguousarray(m["lightData"])
        sc["params"] = json.loads(bytes(m["params_json"]).decode())
    if name == "furnace_cfg3":
        sc["params"] = dict(sc["params"], sun_Power="0", IBL_Power="1.0")
    _cache[name] = sc
    return sc


def load_ibl(name="preview"):
    """preview (the 600x300 map of the reference checkout) | grey (uniform) | 8k (8192x4096 stand-in)"""
    if name == "grey":
        return np.full((8, 16, 4), 128, np.uint8)
    if name == "8k":
        if "ibl_8k" not in _cache:
            _cache["ibl_8k"] = _ibl_8k()
        return _cache["ibl_8k"]
    key = "ibl_" + name
    if key not in _cache:
        _cache[key] = np.ascontiguousarray(np.load(os.path.join(GOLDEN, "ibl_preview.npz"))["rgba"])
    return _cache[key]


def _ibl_8k():
    """8192x4096 stand-in for the missing IBL/Arches_E_PineTree_8k.jpg: bilinear upscale of the 600x300 preview."""
    src = load_ibl("preview").astype(np.float32)
    h, w = src.shape[:2]
    H, W = 4096, 8192
    y = (np.arange(H) + 0.5) * h / H - 0.5
    x = (np.arange(W) + 0.5) * w / W - 0.5
    y0 = np.clip(np.floor(y).astype(int), 0, h - 1); y1 = np.clip(y0 + 1, 0, h - 1); fy = np.clip(y - y0, 0, 1)[:, None, None]
    x0 = np.clip(np.floor(x).astype(int), 0, w - 1); x1 = np.clip(x0 + 1, 0, w - 1); fx = np.clip(x - x0, 0, 1)[None, :, None]
    top = src[y0][:, x0] * (1 - fx) + src[y0][:, x1] * fx
    bot = src[y1][:, x0] * (1 - fx) + src[y1][:, x1] * fx
    return np.ascontiguousarray(np.clip(np.rint(top * (1 - fy) + bot * fy), 0, 255).astype(np.uint8))


def golden():
    if "golden" not in _cache:
        z = np.load(os.path.join(GOLDEN, "golden_ref.npz"))
        _cache["golden"] = {k: z[k] for k in z.files}
    return _cache["golden"]


def cam_env(params, width, height=None):
    """cam[10] / envData[5] as main.py:59-61,72-73 builds them (height only fills cam[7])."""
    height = width if height is None else height
    cam = np.array([float(params["cam_x"]), float(params["cam_y"]), float(params["cam_z"]),
                    float(params["cam_rx"]), float(params["cam_ry"]), float(params["cam_rz"]),
                    width, height, 1, float(params["cam_DOF"]) * (3.14 / 180)]).astype(np.float32)
    env = np.array([float(params["sun_rx"]), float(params["sun_ry"]), float(params["sun_rz"]),
                    float(params["sun_Power"]), float(params["IBL_Power"])]).astype(np.float32)
    return cam, env


def upload(ctx, sc, ibl=None):
    ctx.set_scene(sc["V_p"], sc["V_n"], sc["V_uv"], sc["faceData"], sc["materialData"], sc["lightData"], sc["BVH"])
    ctx.set_ibl(load_ibl() if ibl is None else ibl)
