"""Drop-in replacement for the reference's BVH.py (class BVH, attribute exportArray).

`BVH(faceData, V_p).exportArray` (reference BVH.py:122-144,163-167; used at FileManager.py:245 and
main.py:84-85) is produced by the native builder behind `b200rt_build_bvh` — identical array, node for
node, in milliseconds instead of seconds (the reference's README names the Python build as its main
bottleneck, README.md:28).  The per-node Python objects of the reference (`root`, `nodeList`) only
feed its interactive matplotlib debug viewer (FileManager.py:120-198) and are not reproduced.
"""
import numpy as np

from . import _capi


class BVH(object):

    def __init__(self, faceData, V_p):
        self.V_p = V_p
        self.exportArray, self.depth = _capi.build_bvh(np.asarray(faceData), np.asarray(V_p), return_depth=True)
        self.NodeCounter = self.exportArray.size // 9
