"""Stand-in for `pyopencl` on hosts that have no OpenCL runtime.

The reference's driver (main.py:1,21-26) imports pyopencl and calls exactly four names before it
hands everything to KernelLauncher: get_platforms(), Platform.get_devices(), Context(),
CommandQueue(context).  It indexes platform[1] and platform[0], so two platforms are reported.
The drop-in KernelLauncher ignores all of these objects; this module performs no computation and
is used only when put on PYTHONPATH explicitly."""


class Device(object):
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return "<pyopencl stand-in Device %r>" % (self.name,)


class Platform(object):
    def __init__(self, name, devices):
        self.name = name
        self._devices = devices

    def get_devices(self, device_type=None):
        return list(self._devices)


def get_platforms():
    return [Platform("b200rt (CUDA, sm_100a)", [Device("NVIDIA B200 via libb200rt.so")]),
            Platform("b200rt host placeholder", [Device("placeholder: kernels never run on the host")])]


class Context(object):
    def __init__(self, devices=None, properties=None, dev_type=None):
        self.devices = devices


class CommandQueue(object):
    def __init__(self, context, device=None, properties=None):
        self.context = context

    def finish(self):
        pass
