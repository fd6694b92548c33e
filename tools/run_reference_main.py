#!/usr/bin/env python3
"""Runs the REFERENCE's own driver — main.main(scene), main.py:14-111, the function UI.render calls (UI.py:100) —
unchanged, on this repo's drop-ins:

  * `KernelLauncher.py` and `BVH.py` next to it are the one-line shims of INTEGRATION.md §2 / §4a;
  * `compat/` supplies pyopencl / pywavefront / matplotlib where the host lacks them;
  * the work directory holds `IBL/Arches_E_PineTree_8k.jpg` (main.py:68 opens exactly that name; the 8k JPEG is missing
    upstream, so the preview that IS in the checkout is stored under it) and `output/` (main.py:101-104).

Nothing of the reference is copied or modified: its main.py / FileManager.py / configReader.py are imported from the
checkout, its OBJ + INI are copied into the work directory only so that resolution / spp can be set for the run.

usage: run_reference_main.py <reference_root> <work_dir> [scene="Cornell box"] [resolution=128] [spp=8] [cuda_devices=0]
Leaves <work_dir>/output/out.png and prints one JSON line (frame size, seconds, stats of the launch)."""
import importlib
import json
import os
import shutil
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "ensem3a_openclraytracer_b200", "compat")

SHIM_KL = "from ensem3a_openclraytracer_b200.KernelLauncher import KernelLauncher  # noqa: F401\n"
SHIM_BVH = "from ensem3a_openclraytracer_b200.BVH import BVH, Box, Node  # noqa: F401\n"


def prepare(ref, work, scene, resolution, spp):
    os.makedirs(os.path.join(work, "IBL"), exist_ok=True)
    os.makedirs(os.path.join(work, "output"), exist_ok=True)
    os.makedirs(os.path.join(work, "shims"), exist_ok=True)
    os.makedirs(os.path.join(work, "ObjFiles"), exist_ok=True)
    with open(os.path.join(work, "shims", "KernelLauncher.py"), "w") as f:
        f.write(SHIM_KL)
    with open(os.path.join(work, "shims", "BVH.py"), "w") as f:
        f.write(SHIM_BVH)
    shutil.copyfile(os.path.join(ref, "IBL", "Arches_E_PineTree_Preview.jpg"),
                    os.path.join(work, "IBL", "Arches_E_PineTree_8k.jpg"))
    obj = os.path.join(work, "ObjFiles", scene + ".obj")
    shutil.copyfile(os.path.join(ref, "ObjFiles", scene + ".obj"), obj)
    ini_lines = open(os.path.join(ref, "ObjFiles", scene + ".ini")).read().splitlines()
    out = []
    for line in ini_lines:                     # the .ini is what the UI edits (UI.py:215): set the frame for this run
        key = line.split("=", 1)[0]
        if key == "resolution":
            line = f"resolution={resolution}"
        elif key == "spp":
            line = f"spp={spp}"
        out.append(line)
    with open(obj[:-4] + ".ini", "w") as f:
        f.write("\n".join(out) + "\n")
    return obj


def run(ref, work, scene="Cornell box", resolution=128, spp=8, cuda_devices=(0,)):
    ref, work = os.path.abspath(ref), os.path.abspath(work)
    obj = prepare(ref, work, scene, resolution, spp)
    try:
        import pyopencl  # noqa: F401  (a real PyOpenCL wins when the host has one)
        extra = []
    except ImportError:
        extra = [COMPAT]
    for name in ("pywavefront", "matplotlib"):
        try:
            importlib.import_module(name)
        except ImportError:
            if COMPAT not in extra:
                extra.append(COMPAT)
    sys.path[:0] = [os.path.join(work, "shims"), ROOT] + extra + [ref]
    sys.dont_write_bytecode = True             # the checkout is read-only
    if len(cuda_devices) > 1:                  # main.py:28 passes no such argument: the launcher reads it from here
        os.environ["B200RT_CUDA_DEVICES"] = ",".join(str(d) for d in cuda_devices)
    cwd = os.getcwd()
    os.chdir(work)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_main = importlib.import_module("main")           # the reference's main.py, unmodified
            fm = importlib.import_module("FileManager")
        assert os.path.abspath(ref_main.__file__).startswith(ref), ref_main.__file__
        scene_obj = fm.Scene(obj, True, None)                    # what UI.py:98 does before calling main(scene)
        t0 = time.perf_counter()
        ref_main.main(scene_obj)                                 # main.py:14
        dt = time.perf_counter() - t0
    finally:
        os.chdir(cwd)
    info = {"scene": scene, "resolution": resolution, "spp": spp, "seconds_in_main": dt,
            "out_png": os.path.join(work, "output", "out.png"), "launcher_module": ref_main.KernelLauncher.__module__,
            "bvh_module": type(scene_obj.BVH).__module__, "cuda_devices": list(cuda_devices)}
    return info, scene_obj


if __name__ == "__main__":
    a = sys.argv[1:]
    if len(a) < 2:
        sys.exit(__doc__)
    devs = tuple(int(x) for x in a[5].split(",")) if len(a) > 5 else (0,)
    info, _ = run(a[0], a[1], a[2] if len(a) > 2 else "Cornell box", int(a[3]) if len(a) > 3 else 128,
                  int(a[4]) if len(a) > 4 else 8, devs)
    print(json.dumps(info))
