"""CPU: the C-ABI shared library loads, exports every symbol include/b200rt.h declares, agrees with the
ctypes struct layouts, and fails loudly (no fallback) when no GPU is usable."""
import ctypes
import inspect
import os
import re

import pytest

import ensem3a_openclraytracer_b200 as rt
from ensem3a_openclraytracer_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200rt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rt_[a-z_0-9]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(rt.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.commonpath([ROOT, rt.LIB_PATH]) == ROOT


def test_every_declared_symbol_is_exported_and_bound():
    lib = rt.load_library()
    declared = header_symbols()
    assert declared, "no declarations parsed from include/b200rt.h"
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in b200rt.h but not exported by libb200rt.so"
    assert sorted(_capi.SYMBOLS) == declared, "ctypes binding list and header disagree"


def test_default_opts_and_struct_layout():
    lib = rt.load_library()
    o = rt.Opts()
    ctypes.memset(ctypes.byref(o), 0xAB, ctypes.sizeof(o))
    lib.b200rt_default_opts(ctypes.byref(o))
    assert (o.rng_mode, o.traversal, o.stack_cap, o.output) == (rt.RNG_REFERENCE, rt.TRAVERSAL_FAST, 20, rt.OUT_FINAL)
    assert (o.sample_begin, o.sample_end, o.pixel_begin, o.pixel_end, o.seed, o.collect_stats) == (0, 0, 0, 0, 0, 0)
    assert ctypes.sizeof(rt.Opts) == 64       # 8 x int32, uint64, 6 x int32
    assert ctypes.sizeof(rt.Stats) == 120     # 5 x uint64, 4 x float, 6 x int32, 3 x float, 4 x int32, uint64


def test_version_string():
    assert b"sm_100a" in rt.load_library().b200rt_version()


def test_kernel_launcher_keeps_the_reference_signatures():
    """reference KernelLauncher.py:8, :33, :90 — same names, same positional order."""
    KL = rt.KernelLauncher
    assert list(inspect.signature(KL.__init__).parameters)[:5] == ["self", "context", "platform", "device", "queue"]
    assert list(inspect.signature(KL.launch_Raytracing).parameters) == [
        "self", "h_img_out", "h_vertex_p", "h_vertex_n", "h_vertex_uv", "h_face_data", "h_material_data",
        "h_light_data", "h_BVH", "h_cam", "h_envData", "imgDim", "spp", "maxBounce", "h_IBL"]
    assert list(inspect.signature(KL.launch_ImgProcessing).parameters) == ["self", "h_src", "h_out", "SIZE"]


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="needs a machine WITHOUT a GPU")
def test_no_gpu_is_an_error_not_a_fallback():
    with pytest.raises(rt.B200RTError, match="no CPU fallback"):
        rt.Context(0)
    with pytest.raises(rt.B200RTError):
        rt.KernelLauncher(None, None, None, None)


def test_missing_library_is_an_error(monkeypatch):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", os.path.join(ROOT, "does", "not", "exist.so"))
    with pytest.raises(rt.B200RTError, match="no fallback"):
        _capi.load_library()


def test_tools_do_not_import_the_oracle():
    """Development tools under tools/ drive the product only; anything that needs the checker lives in tests/."""
    tdir = os.path.join(ROOT, "tools")
    for f in sorted(os.listdir(tdir)):
        if f.endswith(".py"):
            text = open(os.path.join(tdir, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"tools/{f} imports oracle"


def test_product_does_not_import_the_oracle():
    """Nothing under the package may reference oracle/ (the checker is never the product)."""
    pkg = os.path.join(ROOT, "ensem3a_openclraytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle"
                assert "rt_oracle" not in text and "libclref" not in text, f"{f} mentions an oracle library"
