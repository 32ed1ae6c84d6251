"""Texture payloads: `RayTracerTexture.ColorData` (RayTracerTypeLibrary/RayTracerTexture.cs:24-33).

The reference decodes textures with GDI+ (`new Bitmap(path)` + `LockBits(Format32bppPArgb)`),
which yields 32-bit 0xAARRGGBB texels in top-down row order whatever the file's own row order.
This module restates that for uncompressed 24/32-bpp BMP files (the only format on the path:
Free_crate/Diffuse.bmp is 512x512 24bpp bottom-up) and provides the deterministic synthetic
512x512 crate texture the benchmarks use (the reference's own BMP does not travel to the GPU box).
"""
import struct

import numpy as np


def read_bmp_argb(path):
    """Decode an uncompressed 24/32-bpp BMP to a (H, W) uint32 array of 0xAARRGGBB, top-down."""
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:2] != b"BM":
        raise ValueError("not a BMP file: %s" % path)
    off_bits = struct.unpack_from("<I", data, 10)[0]
    hdr_size, width, height, planes, bpp, compression = struct.unpack_from("<IiiHHI", data, 14)
    if compression not in (0, 3) or bpp not in (24, 32):
        raise ValueError("unsupported BMP (bpp=%d, compression=%d)" % (bpp, compression))
    bottom_up = height > 0
    height = abs(height)
    stride = ((width * bpp + 31) // 32) * 4
    rows = np.frombuffer(data, dtype=np.uint8, count=stride * height, offset=off_bits).reshape(height, stride)
    px = rows[:, : width * (bpp // 8)].reshape(height, width, bpp // 8).astype(np.uint32)
    b, g, r = px[..., 0], px[..., 1], px[..., 2]
    a = px[..., 3] if bpp == 32 else np.full_like(b, 255)
    if bpp == 32 and not a.any():
        a = np.full_like(b, 255)  # GDI+ treats BI_RGB 32bpp as opaque
    argb = (a << 24) | (r << 16) | (g << 8) | b
    if bottom_up:
        argb = argb[::-1]
    return np.ascontiguousarray(argb, dtype=np.uint32)


def write_bmp24(path, argb):
    """Write a (H, W) 0xAARRGGBB top-down array as a bottom-up 24-bpp BMP (alpha dropped)."""
    argb = np.asarray(argb, dtype=np.uint32)
    h, w = argb.shape
    stride = ((w * 24 + 31) // 32) * 4
    rows = np.zeros((h, stride), dtype=np.uint8)
    px = rows[:, : w * 3].reshape(h, w, 3)
    px[..., 0] = argb & 255
    px[..., 1] = (argb >> 8) & 255
    px[..., 2] = (argb >> 16) & 255
    rows = rows[::-1]
    with open(path, "wb") as fh:
        fh.write(b"BM" + struct.pack("<IHHI", 54 + stride * h, 0, 0, 54))
        fh.write(struct.pack("<IiiHHIIiiII", 40, w, h, 1, 24, 0, stride * h, 2835, 2835, 0, 0))
        fh.write(rows.tobytes())


def synthetic_crate_texture(size=512, seed=7):
    """Deterministic stand-in for Free_crate/Diffuse.bmp: planks, a frame and per-texel grain."""
    rng = np.random.Generator(np.random.PCG64(seed))
    y, x = np.mgrid[0:size, 0:size]
    plank = (x // (size // 8)) % 2
    base_r = 150 + 30 * plank
    base_g = 100 + 22 * plank
    base_b = 50 + 12 * plank
    grain = rng.integers(-18, 19, size=(size, size))
    streak = (12 * np.sin(y * 0.11 + plank * 1.7)).astype(np.int64)
    border = (x < size // 16) | (x >= size - size // 16) | (y < size // 16) | (y >= size - size // 16)
    diag = np.abs(x - y) < size // 40
    r = np.clip(base_r + grain + streak - 60 * border - 35 * diag, 0, 255).astype(np.uint32)
    g = np.clip(base_g + grain + streak - 45 * border - 25 * diag, 0, 255).astype(np.uint32)
    b = np.clip(base_b + grain // 2 + streak - 25 * border - 10 * diag, 0, 255).astype(np.uint32)
    return np.ascontiguousarray((np.uint32(255) << 24) | (r << 16) | (g << 8) | b, dtype=np.uint32)
