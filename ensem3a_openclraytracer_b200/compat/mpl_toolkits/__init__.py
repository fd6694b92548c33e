"""Import-only stand-in, see compat/matplotlib/__init__.py."""
