"""Drop-in replacement for the reference's BVH.py (class BVH, attribute exportArray).

`BVH(faceData, V_p).exportArray` (reference BVH.py:122-144,163-167; used at FileManager.py:245 and
main.py:84-85) is produced by the native builder behind `b200rt_build_bvh` — identical array, node for
node, in milliseconds instead of seconds (the reference's README names the Python build as its main
bottleneck, README.md:28).  The per-node Python objects of the reference (`root`, `nodeList`; BVH.py:5-27) only
feed its interactive matplotlib debug viewer (FileManager.py:107-116); they are rebuilt from the array on first
access, so a caller that never opens the viewer never pays for them.
"""
import numpy as np

from . import _capi


class Box(object):
    """reference BVH.py:5-9"""

    def __init__(self, min, max):
        self.min = min
        self.max = max


class Node(object):
    """What the reference's viewer reads of a node (BVH.py:12-27): `box`, `childL`, `childR` and `array`, which is
    empty for an interior node and holds the leaf's one triangle otherwise (faceData row + triangle id, so that
    `array[0][10]` is the id, FileManager.py:115)."""

    def __init__(self, box, childL, childR, array):
        self.box = box
        self.childL = childL
        self.childR = childR
        self.array = array


class BVH(object):

    def __init__(self, faceData, V_p):
        self.V_p = V_p
        self._face = np.asarray(faceData)
        self.exportArray, self.depth = _capi.build_bvh(self._face, np.asarray(V_p), return_depth=True)
        self.NodeCounter = self.exportArray.size // 9
        self._nodes = None

    @property
    def nodeList(self):
        if self._nodes is None:
            rec = self.exportArray.reshape(-1, 9)
            face = self._face.reshape(-1, 10)
            nodes = []
            for r in rec:
                tri = int(r[8])
                arr = [] if tri == -1 else [list(face[tri]) + [tri]]
                nodes.append(Node(Box(r[2:5].reshape(3, 1).copy(), r[5:8].reshape(3, 1).copy()), int(r[0]), int(r[1]), arr))
            self._nodes = nodes
        return self._nodes

    @property
    def root(self):
        return self.nodeList[0]
