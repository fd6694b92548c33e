"""Import-only stand-in, see compat/matplotlib/__init__.py."""


class Poly3DCollection(object):
    pass


class Line3DCollection(object):
    pass
