"""TEST INFRASTRUCTURE — ctypes access to oracle/_ref/libclref*.so (the reference's own
OpenCL-C text compiled by g++, see oracle/build_ref.py).  Not imported by the product."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def available(flavour="libclref.so"):
    return os.path.exists(os.path.join(_HERE, "_ref", flavour))


def _lib(flavour):
    if flavour not in _LIBS:
        lib = ctypes.CDLL(os.path.join(_HERE, "_ref", flavour))
        lib.ref_raytrace.restype = None
        lib.ref_raytrace.argtypes = [_f32p, _f32p, _f32p, _f32p, _i32p, _i32p, _f32p, _f32p, _f32p, _f32p,
                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, _u8p, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_void_p]
        lib.ref_primary.restype = None
        lib.ref_primary.argtypes = [_f32p, _f32p, _f32p, _i32p, _f32p, _f32p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, _f32p, _i32p, _f32p, _i32p, _i32p, _f32p]
        lib.ref_img_processing.restype = None
        lib.ref_img_processing.argtypes = [_f32p, _f32p, ctypes.c_int, ctypes.c_int]
        lib.ref_rand_stream.restype = None
        lib.ref_rand_stream.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _f32p]
        _LIBS[flavour] = lib
    return _LIBS[flavour]


def _light(scene):
    light = np.ascontiguousarray(scene["lightData"], dtype=np.int32)
    if light.size == 0:  # KernelLauncher.py:64-66 substitutes np.array([0])
        light = np.zeros(1, dtype=np.int32)
    return light


def raytrace(scene, cam, env, img_dim, spp, max_bounce, ibl_rgba, i0=0, i1=None, nthreads=0,
             counters=False, flavour=None):
    """The reference kernel over work-items [i0,i1) of an `img_dim`-pixel launch.
    Returns (out[img_dim*3] float32 (only [i0,i1) filled), counters dict or None)."""
    flavour = flavour or ("libclref_count.so" if counters else "libclref.so")
    lib = _lib(flavour)
    i1 = img_dim if i1 is None else i1
    out = np.zeros(img_dim * 3, dtype=np.float32)
    ibl = np.ascontiguousarray(ibl_rgba, dtype=np.uint8)
    h, w = ibl.shape[0], ibl.shape[1]
    cnt = (ctypes.c_ulonglong * 4)()
    face = scene["faceData"]
    lib.ref_raytrace(out, scene["V_p"], scene["V_n"], scene["V_uv"], face, _light(scene),
                     scene["materialData"], scene["BVH"], np.ascontiguousarray(cam, np.float32),
                     np.ascontiguousarray(env, np.float32), face.size // 10, int(scene["lightData"].size),
                     img_dim, spp, max_bounce, w, h, ibl.reshape(-1), i0, i1, nthreads,
                     ctypes.cast(cnt, ctypes.c_void_p))
    c = None
    if counters:
        c = {"rays": cnt[0], "box_tests": cnt[1], "tri_tests": cnt[2], "rand_calls": cnt[3]}
    return out, c


def primary(scene, cam, img_dim, i0=0, i1=None):
    lib = _lib("libclref.so")
    i1 = img_dim if i1 is None else i1
    n = i1 - i0
    d = np.zeros(n * 3, np.float32)
    hit = np.zeros(n, np.int32)
    k = np.zeros(n, np.float32)
    mat = np.zeros(n, np.int32)
    tri = np.zeros(n, np.int32)
    nrm = np.zeros(n * 3, np.float32)
    face = scene["faceData"]
    lib.ref_primary(scene["V_p"], scene["V_n"], scene["V_uv"], face, scene["BVH"],
                    np.ascontiguousarray(cam, np.float32), face.size // 10, i0, i1, d, hit, k, mat, tri, nrm)
    return {"dir": d.reshape(n, 3), "hit": hit, "k": k, "mat": mat, "tri": tri, "n": nrm.reshape(n, 3)}


def img_processing(src, n, global_size=None):
    lib = _lib("libclref.so")
    src = np.ascontiguousarray(src, np.float32)
    out = np.zeros_like(src)
    lib.ref_img_processing(src, out, n, src.size if global_size is None else global_size)
    return out


def rand_stream(pixel, img_size, n):
    lib = _lib("libclref.so")
    out = np.zeros(n, np.float32)
    lib.ref_rand_stream(pixel, img_size, n, out)
    return out
