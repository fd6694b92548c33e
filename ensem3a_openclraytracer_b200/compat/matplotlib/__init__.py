"""Import-only stand-in for `matplotlib`: the reference's FileManager.py imports pyplot and mplot3d at module level
(FileManager.py:10-13) but only its interactive BVH debug viewer (Scene.test, call commented out at :250) uses
them.  Any attempt to actually plot raises."""
