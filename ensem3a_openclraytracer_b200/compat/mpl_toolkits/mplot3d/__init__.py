"""Import-only stand-in, see compat/matplotlib/__init__.py."""


class Axes3D(object):
    pass
