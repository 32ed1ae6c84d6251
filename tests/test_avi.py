"""Frame sink (xna_ray_trace_b200/avi.py): the AviFile wrapper's API on a Linux RIFF/OpenDML writer.
CPU only.  Round trips (write -> read back), header conventions of the wrapper
(aviFileWrapper_src/VideoStream.cs:229-271, 344-384) and multi-segment files."""
import struct

import numpy as np
import pytest

from xna_ray_trace_b200 import avi


def frames(n, h, w, seed=5):
    rng = np.random.default_rng(seed)
    return [(rng.integers(0, 1 << 24, size=(h, w), dtype=np.uint32) | np.uint32(0xFF000000)) for _ in range(n)]


def test_get_rate_and_scale():
    # VideoStream.GetRateAndScale: multiply by 10 until integral
    assert avi.get_rate_and_scale(30) == (30, 1)
    assert avi.get_rate_and_scale(29.97) == (2997, 100)
    assert avi.get_rate_and_scale(12.5) == (125, 10)


def test_round_trip_bitmap_overload(tmp_path):
    fr = frames(5, 18, 23)   # odd width: 24 bpp rows are padded to 4 bytes (Stride * Height)
    path = str(tmp_path / "a.avi")
    m = avi.AviManager(path, False)
    s = m.AddVideoStream(True, 30, fr[0])       # Game1.cs:198: compressed requested, written uncompressed here
    assert s.WriteCompressed is False and s.CountFrames == 1
    for f in fr[1:]:
        s.AddFrame(f)
    m.Close()
    r = avi.AviManager(path, True)
    v = r.GetVideoStream()
    assert (v.CountFrames, v.Width, v.Height, v.CountBitsPerPixel, v.FrameRate) == (5, 23, 18, 24, 30.0)
    for i, f in enumerate(fr):
        assert np.array_equal(v.GetBitmap(i), f)
    assert r.StreamInfo["fccType"] == b"vids" and r.StreamInfo["dwQuality"] == -1
    assert (r.StreamInfo["dwScale"], r.StreamInfo["dwRate"], r.StreamInfo["dwLength"]) == (1, 30, 5)
    assert r.StreamInfo["rcFrame"] == (0, 0, 23, 18)
    assert r.MainHeader["dwMicroSecPerFrame"] == 33333 and r.MainHeader["dwWidth"] == 23
    # on disk: bottom-up BGR rows (AddFrame flips with RotateNoneFlipY before AVIStreamWrite)
    raw = open(path, "rb").read()
    pos = raw.index(b"00db", raw.index(b"movi"))  # the first frame chunk (the super index also names the stream)
    size = struct.unpack("<I", raw[pos + 4:pos + 8])[0]
    stride = (23 * 3 + 3) & ~3
    assert size == stride * 18
    first_row = np.frombuffer(raw[pos + 8:pos + 8 + 3], dtype=np.uint8)
    px = int(fr[0][17, 0])  # bottom row first
    assert tuple(first_row) == ((px >> 16) & 255, (px >> 8) & 255, px & 255)
    assert raw[:4] == b"RIFF" and raw[8:12] == b"AVI " and b"idx1" in raw and b"indx" in raw


def test_explicit_format_overload_and_rgb_arrays(tmp_path):
    path = str(tmp_path / "b.avi")
    m = avi.AviManager(path, False)
    s = m.AddVideoStream(False, 12.5, 0, 16, 8, avi.PIXELFORMAT_32BPP_ARGB)
    rgb = np.zeros((8, 16, 3), dtype=np.uint8)
    rgb[..., 0], rgb[..., 1], rgb[..., 2] = 10, 20, 30
    s.AddFrame(rgb)
    with pytest.raises(ValueError):
        s.AddFrame(np.zeros((9, 16), dtype=np.uint32))
    m.Close()
    r = avi.AviManager(path, True)
    v = r.GetVideoStream()
    assert (r.StreamInfo["dwScale"], r.StreamInfo["dwRate"]) == (10, 125)
    assert v.CountBitsPerPixel == 32 and v.CountFrames == 1
    assert (v.GetBitmap(0) == np.uint32(10 | 20 << 8 | 30 << 16 | 255 << 24)).all()
    with pytest.raises(IOError):
        avi.AviManager(str(tmp_path / "missing.avi"), True)


def test_opendml_segments(tmp_path):
    """More data than one RIFF segment holds: 'AVIX' segments + super index, all frames readable."""
    fr = frames(40, 32, 32, seed=9)      # 3 KiB per frame
    path = str(tmp_path / "c.avi")
    m = avi.AviManager(path, False, first_segment_limit=20_000, segment_limit=30_000)
    s = m.AddVideoStream(False, 30, 0, 32, 32, avi.PIXELFORMAT_24BPP_RGB)
    for f in fr:
        s.AddFrame(f)
    m.Close()
    raw = open(path, "rb").read()
    assert raw.count(b"AVIX") >= 3
    r = avi.AviManager(path, True)
    v = r.GetVideoStream()
    assert v.CountFrames == 40 and len(r.Segments) >= 4
    assert sum(d for _, _, d in r.Segments) == 40
    assert r.MainHeader["dwTotalFrames"] < 40          # AVI 1.0 view: the first RIFF only
    for i in (0, 5, 17, 39):
        assert np.array_equal(v.GetBitmap(i), fr[i])
    # every RIFF size field is consistent with the file
    pos, n = 0, 0
    while pos < len(raw):
        assert raw[pos:pos + 4] == b"RIFF"
        pos += 8 + struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        n += 1
    assert pos == len(raw) and n == len(r.Segments)


def test_threaded_payload_writes_give_the_same_file(tmp_path):
    """AviManager(io_threads=n): payloads written by worker threads at their reserved offsets; the file is byte for byte
    the one the sequential writer produces, across RIFF segment boundaries (AVIX) and odd payload sizes."""
    rng = np.random.default_rng(2)
    frames = [rng.integers(0, 2 ** 32, size=(9, 13), dtype=np.uint64).astype(np.uint32) for _ in range(40)]
    paths = []
    for threads in (0, 3):
        p = str(tmp_path / ("t%d.avi" % threads))
        m = avi.AviManager(p, False, first_segment_limit=4000, segment_limit=3000, io_threads=threads)
        s = m.AddVideoStream(False, 30, 0, 13, 9, avi.PIXELFORMAT_24BPP_RGB)
        for k, f in enumerate(frames):
            if k % 2:
                bits, _ = avi.to_dib_bits(f, 24)
                s.AddFrameBits(bits)
            else:
                s.AddFrame(f)
            if k == 17:
                s.Flush()
        m.Close()
        paths.append(p)
    a, b = open(paths[0], "rb").read(), open(paths[1], "rb").read()
    assert a == b and len(a) > 40 * 9 * 40
    v = avi.AviManager(paths[1], True).GetVideoStream()
    assert v.CountFrames == 40 and np.array_equal(v.GetBitmap(33) & 0xFFFFFF, frames[33] & 0xFFFFFF)
