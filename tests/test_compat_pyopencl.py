"""The pyopencl stand-in offers exactly what the reference's main.py:21-26 calls."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    p = os.path.join(ROOT, "ensem3a_openclraytracer_b200", "compat", "pyopencl", "__init__.py")
    spec = importlib.util.spec_from_file_location("pyopencl_standin", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_driver_preamble_runs():
    cl = _load()
    platform = cl.get_platforms()
    cpu = platform[1].get_devices()      # main.py:22
    gpu = platform[0].get_devices()      # main.py:23
    context = cl.Context()               # main.py:24
    queue = cl.CommandQueue(context)     # main.py:26
    assert len(cpu) >= 1 and len(gpu) >= 1 and queue.context is context


def test_stand_in_has_no_compute_surface():
    cl = _load()
    for name in ("Program", "Buffer", "Image", "enqueue_copy", "mem_flags"):
        assert not hasattr(cl, name)
