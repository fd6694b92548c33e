#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — builds oracle/_ref/ from the reference's own kernel sources.

The reference hot path is OpenCL C (Kernels/Raytracing.cl, MathLib.cl, stack.cl,
ImgProcessing.cl).  There is no OpenCL runtime in this image, so the sources are
compiled as host C++ by g++ where they lie:

  1. each .cl is read from <reference>/Kernels/ and written to a TEMPORARY directory with two
     mechanical regex rewrites (no statement is added, removed or reordered):
        (floatN)(...) / (int2)(...)   ->  floatN(...) / int2(...)   vector literal -> ctor
        .yzw / .xyz                    ->  .yzw() / .xyz()            swizzle -> accessor
     plus, for the counting build only, a `CLREF_COUNT(x);` statement inserted as the
     first statement of rayTrace / interNode / intersect / rand;
  2. oracle/ref_shim/ref_driver.cpp (#include "cl_shim.h", #include "Raytracing.cl")
     is compiled twice:  libclref.so (timing build) and libclref_count.so (counters).

oracle/_ref/ holds only the resulting shared objects.  It is git-ignored but travels to the GPU box
with the gpurun snapshot, where /root/reference does not exist; no reference text is stored anywhere.
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SHIM = os.path.join(HERE, "ref_shim")
FILES = ["stack.cl", "MathLib.cl", "Raytracing.cl", "ImgProcessing.cl"]
COUNTED = {  # function name -> counter field
    "rayTrace": "rays",
    "interNode": "box_tests",
    "intersect": "tri_tests",
    "rand": "rand_calls",
}
CXXFLAGS = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-std=c++17",
            "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-maybe-uninitialized",
            "-Wno-uninitialized"]


def transpile(text: str) -> str:
    text = re.sub(r"\((float[234]|int2)\)\s*\(", r"\1(", text)
    text = re.sub(r"\.(yzw|xyz)\b", r".\1()", text)
    for fn, field in COUNTED.items():
        # definition = "<ret type> fn(<params>)\s*{"  (params never contain braces or ';')
        pat = re.compile(r"(\b(?:hitInfo|bool|float)\s+" + fn + r"\s*\([^;{}]*\)\s*\{)")
        text, n = pat.subn(r"\1 CLREF_COUNT(" + field + ");", text, count=1)
    return text


def build(reference_root: str = "/root/reference", verbose: bool = True) -> bool:
    """Returns True when oracle/_ref/libclref*.so exist afterwards."""
    libs = [os.path.join(OUT, "libclref.so"), os.path.join(OUT, "libclref_count.so")]
    kdir = os.path.join(reference_root, "Kernels")
    if not os.path.isdir(kdir):
        ok = all(os.path.exists(p) for p in libs)
        if verbose:
            print(f"[build_ref] {kdir} absent; prebuilt oracle/_ref libs {'found' if ok else 'MISSING'}")
        return ok
    os.makedirs(OUT, exist_ok=True)
    for stale in FILES:  # earlier versions of this script left the rewritten text here
        if os.path.exists(os.path.join(OUT, stale)):
            os.remove(os.path.join(OUT, stale))
    # the rewritten kernel text lives only in a temporary directory for the duration of the compile:
    # oracle/_ref/ receives nothing but the shared objects
    with tempfile.TemporaryDirectory(prefix="clref_") as tmp:
        for f in FILES:
            with open(os.path.join(kdir, f), "r") as fh:
                src = fh.read()
            with open(os.path.join(tmp, f), "w") as fh:
                fh.write(transpile(src))
        drv = os.path.join(SHIM, "ref_driver.cpp")
        # timing build, counting build, and a sensitivity build (glibc binary32 transcendentals instead of
        # correctly-rounded ones)
        for lib, extra in ((libs[0], []), (libs[1], ["-DCLREF_COUNTERS"]),
                           (os.path.join(OUT, "libclref_libmf.so"), ["-DCLREF_LIBM_FLOAT"])):
            cmd = ["g++"] + CXXFLAGS + extra + ["-I", SHIM, "-I", tmp, drv, "-o", lib]
            if verbose:
                print("[build_ref]", " ".join(cmd))
            subprocess.check_call(cmd)
    return True


if __name__ == "__main__":
    ok = build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    sys.exit(0 if ok else 1)
