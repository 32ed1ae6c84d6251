"""xna-ray-trace_b200 — B200-native drop-in for the per-pixel trace loop of eitan3/xna-ray-trace.

Only what the path needs lives here: `csrc/` (hand-written CUDA for sm_100a + the C-ABI of
include/xnart.h, built into libxnart.so), `capi` (ctypes binding of that ABI), `host` (mirror of
the reference's host types: Mesh, Material, SceneObject, Camera, RayTracer ...), `scenes`
(synthetic inputs of BASELINE.json's configs) and `xna` (host-side XNA value-type math).
There is no CPU fallback: importing is cheap, but any compute call fails loudly if libxnart.so
is missing or no CUDA device is present.
"""
__version__ = "0.1.0"
