"""Synthetic scene of BASELINE.json config 5: a regular grid height field, z = A sin cos plus a
hashed jitter so that no two triangle centroids coincide (BVH.py does not terminate on coincident centroids),
unique v / vt / vn per vertex, one diffuse material.  Buffers in the reference's layouts."""
import numpy as np


def height_field(quads, seed=0, size=10.0, amp=0.6):
    """(V_p float32[3*nv], faceData int32[10*nt]) of a quads x quads grid = 2*quads^2 triangles."""
    n = quads + 1
    rng = np.random.default_rng(seed)
    u = np.linspace(-0.5, 0.5, n, dtype=np.float64)
    x, y = np.meshgrid(u * size, u * size, indexing="xy")
    jit = rng.uniform(-0.2, 0.2, (2, n, n)) * (size / quads)
    x = x + jit[0]
    y = y + jit[1]
    z = amp * np.sin(3.0 * x) * np.cos(2.0 * y) + rng.uniform(-0.02, 0.02, (n, n))
    vp = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(quads), np.arange(quads), indexing="xy")
    v00 = (j * n + i).ravel()
    v10, v01, v11 = v00 + 1, v00 + n, v00 + n + 1
    tri = np.concatenate([np.stack([v00, v10, v11], 1), np.stack([v00, v11, v01], 1)], axis=0).astype(np.int32)
    face = np.zeros((tri.shape[0], 10), np.int32)          # [mat, uv0..2, n0..2, p0..2]; one vt / vn per vertex
    face[:, 1:4] = tri
    face[:, 4:7] = tri
    face[:, 7:10] = tri
    return vp.reshape(-1), face.reshape(-1)


def height_field_scene(quads, seed=0):
    """All buffers launch_Raytracing needs except the BVH (built by the caller)."""
    vp, face = height_field(quads, seed)
    p = vp.reshape(-1, 3)
    tri = face.reshape(-1, 10)[:, 7:10]
    # per-vertex normal = normalised sum of adjacent face normals (float32, like an OBJ exporter would write)
    fn = np.cross(p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]])
    vn = np.zeros_like(p)
    for k in range(3):
        np.add.at(vn, tri[:, k], fn)
    vn /= np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-20)
    uv = ((p[:, :2] - p[:, :2].min(0)) / np.ptp(p[:, :2], axis=0)).astype(np.float32)
    mat = np.array([1, 0.8, 0.8, 0.8, 0.0, 1.0], np.float32)
    return {"V_p": vp, "V_n": vn.astype(np.float32).reshape(-1), "V_uv": uv.reshape(-1), "faceData": face,
            "materialData": mat, "lightData": np.zeros(0, np.int32)}
