"""Importable alias of the `xna-ray-trace_b200/` package directory.

A hyphen is not a valid Python module name, so this thin package puts the real directory on its
search path: `import xna_ray_trace_b200.capi` loads `xna-ray-trace_b200/capi.py`.
There is no CPU fallback anywhere below: compute calls fail loudly when libxnart.so or a CUDA
device is missing.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "xna-ray-trace_b200"))
__version__ = "0.1.0"
