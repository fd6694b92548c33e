"""See compat/matplotlib/__init__.py."""


def __getattr__(name):
    raise AttributeError("matplotlib stand-in: plotting (pyplot.%s) is not available; install matplotlib for the "
                         "reference's BVH debug viewer" % name)
