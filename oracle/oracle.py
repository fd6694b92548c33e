"""TEST INFRASTRUCTURE — ctypes access to oracle/_build/librt_oracle.so (rt_oracle.c, the CPU
restatement of the reference hot path).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this; the product never does."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


class OrcOpts(ctypes.Structure):
    _fields_ = [("rng_mode", ctypes.c_int), ("seed_lo", ctypes.c_uint32), ("seed_hi", ctypes.c_uint32),
                ("stack_cap", ctypes.c_int), ("s0", ctypes.c_int), ("s1", ctypes.c_int),
                ("raw_sums", ctypes.c_int), ("nthreads", ctypes.c_int), ("sampling", ctypes.c_int),
                ("n_light", ctypes.c_int), ("light", ctypes.c_void_p)]


def build(verbose=False):
    subprocess.check_call(["make", "-C", _HERE] + ([] if verbose else ["-s"]))


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "librt_oracle.so")
        if not os.path.exists(path):
            build()
        lib = ctypes.CDLL(path)
        lib.orc_render.restype = None
        lib.orc_render.argtypes = [_f32p, _f32p, _f32p, _f32p, _i32p, _f32p, _f32p, _f32p, _f32p, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, _u8p, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.POINTER(OrcOpts), ctypes.c_void_p]
        lib.orc_primary.restype = None
        lib.orc_primary.argtypes = [_f32p, _f32p, _f32p, _i32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, _f32p, _i32p, _f32p, _i32p]
        lib.orc_trace_rays.restype = None
        lib.orc_trace_rays.argtypes = [_f32p, _f32p, _f32p, _i32p, _f32p, ctypes.c_int, ctypes.c_int, _f32p,
                                       ctypes.c_int, _i32p, _f32p, ctypes.c_void_p]
        lib.orc_img_processing.restype = None
        lib.orc_img_processing.argtypes = [_f32p, _f32p, ctypes.c_int, ctypes.c_int]
        lib.orc_rand_stream.restype = None
        lib.orc_rand_stream.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                        ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int, _f32p]
        lib.orc_philox.restype = None
        lib.orc_philox.argtypes = [_u32p, ctypes.c_uint32, ctypes.c_uint32, _u32p]
        lib.orc_math_probe.restype = None
        lib.orc_math_probe.argtypes = [ctypes.c_int, _f32p, _f32p, ctypes.c_int, _f32p]
        lib.orc_rotate.restype = None
        lib.orc_rotate.argtypes = [ctypes.c_float, _f32p, _f32p, _f32p]
        _LIB = lib
    return _LIB


RNG_REFERENCE = 0
RNG_PHILOX = 1


def render(scene, cam, env, img_dim, spp, max_bounce, ibl_rgba, i0=0, i1=None, rng_mode=RNG_REFERENCE,
           seed=0, stack_cap=20, s0=0, s1=0, raw_sums=False, nthreads=0, sampling=0):
    """Returns (out[img_dim*3] float32, counters dict)."""
    lib = _lib()
    i1 = img_dim if i1 is None else i1
    out = np.zeros(img_dim * 3, dtype=np.float32)
    ibl = np.ascontiguousarray(ibl_rgba, dtype=np.uint8)
    h, w = ibl.shape[0], ibl.shape[1]
    light = np.ascontiguousarray(scene.get("lightData", np.zeros(0, np.int32)), dtype=np.int32)
    opts = OrcOpts(rng_mode, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, stack_cap, s0, s1,
                   1 if raw_sums else 0, nthreads, sampling, int(light.size),
                   light.ctypes.data_as(ctypes.c_void_p) if light.size else None)
    cnt = (ctypes.c_ulonglong * 4)()
    face = scene["faceData"]
    lib.orc_render(out, scene["V_p"], scene["V_n"], scene["V_uv"], face, scene["materialData"], scene["BVH"],
                   np.ascontiguousarray(cam, np.float32), np.ascontiguousarray(env, np.float32),
                   face.size // 10, img_dim, spp, max_bounce, ibl.reshape(-1), w, h, i0, i1,
                   ctypes.byref(opts), ctypes.cast(cnt, ctypes.c_void_p))
    return out, {"rays": cnt[0], "box_tests": cnt[1], "tri_tests": cnt[2], "rand_calls": cnt[3]}


def primary(scene, cam, img_dim, i0=0, i1=None, stack_cap=20):
    lib = _lib()
    i1 = img_dim if i1 is None else i1
    n = i1 - i0
    d = np.zeros(n * 3, np.float32)
    tri = np.zeros(n, np.int32)
    k = np.zeros(n, np.float32)
    mat = np.zeros(n, np.int32)
    face = scene["faceData"]
    lib.orc_primary(scene["V_p"], scene["V_n"], scene["V_uv"], face, scene["BVH"],
                    np.ascontiguousarray(cam, np.float32), face.size // 10, stack_cap, i0, i1, d, tri, k, mat)
    return {"dir": d.reshape(n, 3), "tri": tri, "k": k, "mat": mat}


def trace_rays(scene, rays, stack_cap=20):
    """rays: (n,6) float32 [ox,oy,oz,dx,dy,dz].  Returns tri (−1 = miss), k, counters."""
    lib = _lib()
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
    n = rays.shape[0]
    tri = np.zeros(n, np.int32)
    k = np.zeros(n, np.float32)
    cnt = (ctypes.c_ulonglong * 4)()
    face = scene["faceData"]
    lib.orc_trace_rays(scene["V_p"], scene["V_n"], scene["V_uv"], face, scene["BVH"], face.size // 10,
                       stack_cap, rays.reshape(-1), n, tri, k, ctypes.cast(cnt, ctypes.c_void_p))
    return tri, k, {"rays": cnt[0], "box_tests": cnt[1], "tri_tests": cnt[2]}


def img_processing(src, n, global_size=None):
    src = np.ascontiguousarray(src, np.float32)
    out = np.zeros_like(src)
    _lib().orc_img_processing(src, out, n, src.size if global_size is None else global_size)
    return out


def rand_stream(mode, pixel, img_size, n, seed=0, sample=0, bounce=0):
    out = np.zeros(n, np.float32)
    _lib().orc_rand_stream(mode, pixel, img_size, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF, sample, bounce,
                           n, out)
    return out


def philox(ctr, k0, k1):
    out = np.zeros(4, np.uint32)
    _lib().orc_philox(np.ascontiguousarray(ctr, np.uint32), k0, k1, out)
    return out


MATH_FN = {"sin": 0, "cos": 1, "acos": 2, "asin": 3, "atan2": 4, "tan": 5, "pow": 6, "div": 7, "sqrt": 8}


def math_probe(fn, a, b=None):
    a = np.ascontiguousarray(a, np.float32)
    b = np.zeros_like(a) if b is None else np.ascontiguousarray(b, np.float32)
    out = np.zeros_like(a)
    _lib().orc_math_probe(MATH_FN[fn], a, b, a.size, out)
    return out


def rotate(angle, axis, vec):
    out = np.zeros(3, np.float32)
    _lib().orc_rotate(float(np.float32(angle)), np.ascontiguousarray(axis, np.float32),
                      np.ascontiguousarray(vec, np.float32), out)
    return out
