#!/usr/bin/env python3
"""Golden BVH of the synthetic height field, produced by the reference's own BVH.py (run in the authoring
container, where /root/reference exists):   python tests/golden/make_bvh_golden.py
Writes tests/golden/bvh_heightfield40.npz = exportArray of BVH(faceData, V_p) for tests.synthetic.height_field(40)."""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
from tests.synthetic import height_field  # noqa: E402


def main(reference="/root/reference"):
    sys.path.insert(0, reference)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from BVH import BVH  # the reference's builder, unchanged
    vp, face = height_field(40, seed=0)
    with contextlib.redirect_stdout(io.StringIO()):   # it prints a progress line per leaf
        b = BVH(face, vp)
    out = os.path.join(HERE, "bvh_heightfield40.npz")
    np.savez_compressed(out, BVH=np.asarray(b.exportArray, np.float32), quads=40, seed=0)
    print("wrote", out, b.exportArray.size // 9, "nodes")


if __name__ == "__main__":
    main(*sys.argv[1:])
