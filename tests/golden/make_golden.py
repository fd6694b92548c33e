#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz.  Runs ONLY in the authoring container (needs /root/reference).

Inputs  — the hot path's input buffers, produced by the reference's own unmodified host code
          (FileManager.Scene + BVH.py, through oracle/ref_scene.py) for every shipped scene,
          and the only environment map present in the checkout decoded the way main.py:68 does
          (PIL .convert("RGBA")).
Outputs — primary hits and rendered pixels of oracle/_ref/libclref.so, i.e. the reference's own
          Kernels/*.cl compiled by g++ (oracle/build_ref.py).  These are the golden vectors the
          CPU restatement (oracle/rt_oracle.c) and the CUDA path are checked against on the GPU
          box, where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_ref, ref_lib, ref_scene, oracle  # noqa: E402

SCENES = {  # fixture name -> reference OBJ stem
    "cornell": "Cornell box",
    "monkey": "Cornell box_Monkey",
    "furnace": "FurnaceHD",
    "serre": "Serre_leger",
    "proto": "protoEnsem",
    "single": "singleTriangle",
}

# the shipped Cornell box_Monkey.ini is all-diffuse; pin a variant with a
# glossy monkey, a glass red wall and an emitting lamp.
MONKEY_CFG2 = {
    "M_4_Type": 2, "M_4_roughness": 0.2,
    "M_1_Type": 3,
    "M_3_Type": 0, "M_3_roughness": 5,
    "IBL_Power": 1.0,
}
# §8d config 3: furnace under a uniform environment, no sun
FURNACE_CFG3 = {"sun_Power": 0, "IBL_Power": 1.0}

PRIMARY_RES = 128
RENDER_RES = 64
RENDER_SPP = 8


def save_scene(name, sc):
    arrays = {k: v for k, v in sc.items() if k != "params"}
    arrays["params_json"] = np.frombuffer(json.dumps(sc["params"]).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, f"scene_{name}.npz"), **arrays)


def main():
    assert build_ref.build(), "oracle/_ref could not be built"
    oracle.build()
    ibl = np.asarray(Image.open(os.path.join(ref_scene.REFERENCE_ROOT, "IBL", "Arches_E_PineTree_Preview.jpg"))
                     .convert("RGBA"))
    np.savez_compressed(os.path.join(HERE, "ibl_preview.npz"), rgba=ibl)

    golden = {}
    for name, stem in SCENES.items():
        sc = ref_scene.load_reference_scene(stem)
        save_scene(name, sc)
        variants = {"": sc}
        if name == "monkey":
            v = ref_scene.load_reference_scene(stem, ini_overrides=MONKEY_CFG2)
            assert np.array_equal(v["BVH"], sc["BVH"])
            np.savez_compressed(os.path.join(HERE, "scene_monkey_cfg2_materials.npz"),
                                materialData=v["materialData"], lightData=v["lightData"],
                                params_json=np.frombuffer(json.dumps(v["params"]).encode(), dtype=np.uint8))
            variants["_cfg2"] = v
        if name == "furnace":
            v = ref_scene.load_reference_scene(stem, ini_overrides=FURNACE_CFG3)
            variants["_cfg3"] = v
        for suffix, s in variants.items():
            key = name + suffix
            # primary hits, square frame
            cam, env = ref_scene.cam_env_from_params(s["params"], PRIMARY_RES)
            p = ref_lib.primary(s, cam, PRIMARY_RES * PRIMARY_RES)
            golden[f"{key}/primary_tri"] = p["tri"]
            golden[f"{key}/primary_k"] = p["k"]
            golden[f"{key}/primary_mat"] = p["mat"].astype(np.int8)
            # full render, reference RNG
            cam, env = ref_scene.cam_env_from_params(s["params"], RENDER_RES)
            img_ibl = ibl if suffix != "_cfg3" else np.full((8, 16, 4), 128, np.uint8)
            out, cnt = ref_lib.raytrace(s, cam, env, RENDER_RES * RENDER_RES, RENDER_SPP, 4, img_ibl, counters=True)
            golden[f"{key}/render"] = out
            golden[f"{key}/counters"] = np.array([cnt["rays"], cnt["box_tests"], cnt["tri_tests"], cnt["rand_calls"]],
                                                 dtype=np.uint64)
            golden[f"{key}/cam"] = cam
            golden[f"{key}/env"] = env
            print(key, "tris", s["faceData"].size // 10, "hit frac", (p["tri"] >= 0).mean(), cnt, flush=True)

    # non-square launch exactly as the reference kernel treats it: cam[6] = width, imgDim = W*H
    s = ref_scene.load_reference_scene(SCENES["cornell"])
    cam, env = ref_scene.cam_env_from_params(s["params"], 96)
    out, _ = ref_lib.raytrace(s, cam, env, 96 * 54, 4, 4, ibl)
    golden["cornell_96x54/render"] = out
    golden["cornell_96x54/cam"] = cam
    golden["cornell_96x54/env"] = env
    p = ref_lib.primary(s, cam, 96 * 54)
    golden["cornell_96x54/primary_tri"] = p["tri"]
    golden["cornell_96x54/primary_k"] = p["k"]

    # rotated camera + rotated sun + maxBounce 2 on a scene with glossy/glass materials
    s = ref_scene.load_reference_scene(SCENES["serre"])
    cam, env = ref_scene.cam_env_from_params(s["params"], 48)
    cam[3:6] = [-30.0, 10.0, 35.0]
    env[0:3] = [20.0, -45.0, 70.0]
    out, _ = ref_lib.raytrace(s, cam, env, 48 * 48, 6, 2, ibl)
    golden["serre_rot/render"] = out
    golden["serre_rot/cam"] = cam
    golden["serre_rot/env"] = env

    # reference RNG stream and tonemap known answers
    for px in (0, 1, 2, 77, 4095, 65535, 262143):
        golden[f"rand/{px}"] = ref_lib.rand_stream(px, 512 * 512, 64)
    x = np.linspace(-0.25, 1.5, 701, dtype=np.float32)
    golden["tonemap/in"] = x
    golden["tonemap/out"] = ref_lib.img_processing(x, 600)  # N < global: the tail stays untouched (0)

    # Philox-mode render from the restatement (the reference has no such mode); guards against drift
    s = ref_scene.load_reference_scene(SCENES["cornell"])
    cam, env = ref_scene.cam_env_from_params(s["params"], RENDER_RES)
    out, cnt = oracle.render(s, cam, env, RENDER_RES * RENDER_RES, RENDER_SPP, 4, ibl, rng_mode=oracle.RNG_PHILOX, seed=7)
    golden["cornell_philox7/render"] = out

    np.savez_compressed(os.path.join(HERE, "golden_ref.npz"), **golden)
    total = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print("fixtures written, total bytes:", total)


if __name__ == "__main__":
    main()
