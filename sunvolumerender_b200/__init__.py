"""sunvolumerender_b200 -- B200-native drop-in for the render hot path of SunVolumeRender.

The product is sunvolumerender_b200/libsvr_b200.so (hand-written CUDA for sm_100a, C ABI in
include/svr_render.h).  This package is the Python host side used by the tests and bench.py:
`_lib` binds the C ABI with ctypes, `scene` builds cameras / transfer functions / lights /
configurations on the host, `render` drives the entry points the way gui/canvas.cpp does.
Importing `render` needs torch; `_lib` and `scene` do not.
"""
from . import _lib, scene  # noqa: F401

__all__ = ["_lib", "scene"]
