"""Frame sink keeping the AviFile wrapper's API on Linux.

The reference's video path (`Game1.compileVideo`, Game1.cs:192-210) drives the third-party AviFile
wrapper: `new AviManager(fileName, false)` (aviFileWrapper_src/AviManager.cs:33),
`AddVideoStream(isCompressed, frameRate, firstFrame)` (:147) /
`AddVideoStream(isCompressed, frameRate, frameSize, width, height, format)` (:124),
`VideoStream.AddFrame(bmp)` (VideoStream.cs:344-365) and `AviManager.Close()` (:327).  That wrapper
is a P/Invoke shell over Windows' avifil32.dll (Avi.cs:175-377), so nothing of it can run here; this
module writes the RIFF container itself and keeps the names, argument meaning and on-disk
conventions of the calls above:

  * stream header as `CreateStreamWithoutFormat` fills it (VideoStream.cs:238-271): fccType 'vids',
    dwScale/dwRate from `GetRateAndScale` (:229-235: multiply by 10 until the rate is integral),
    dwQuality -1, rcFrame = (0, 0, width, height); format = BITMAPINFOHEADER as `SetFormat` (:373-384);
  * frames are stored bottom-up (`RotateNoneFlipY`, :345) as raw DIB bits, rows padded to 4 bytes
    (`bmpDat.Stride * bmpDat.Height`, :354-356), one '00db' chunk per frame;
  * the file is OpenDML (AVI 2.0): 600 frames of 1080p x 24 bpp are 3.7 GB, beyond the 2 GB limit of
    the AVI 1.0 layout avifil32 writes.  The first RIFF 'AVI ' segment carries the legacy 'idx1'
    index, further data goes to RIFF 'AVIX' segments, every segment has an 'ix00' chunk index and the
    stream header a super index ('indx').

Compression: the reference asks for a compressed stream (`AddVideoStream(true, ...)`), which opens a
Windows codec dialog (VideoStream.cs:298); there is no codec here, so the stream is written
uncompressed (BI_RGB) whatever `isCompressed` says and `VideoStream.WriteCompressed` reports False.

A "Bitmap" is a numpy array: (H, W) uint32 packed R | G<<8 | B<<16 | A<<24 (what
`xnart_render_frame` returns, the XNA `Color` layout) or (H, W, 3|4) uint8 RGB(A), top-down rows.

`AviManager(fileName, True)` opens a file written by this module (GetVideoStream, CountFrames,
GetBitmap) — enough for round-trip tests, not a general AVI reader.
"""
import os
import struct

import numpy as np

PIXELFORMAT_24BPP_RGB = 0x21808   # System.Drawing.Imaging.PixelFormat.Format24bppRgb
PIXELFORMAT_32BPP_RGB = 0x22009   # Format32bppRgb
PIXELFORMAT_32BPP_ARGB = 0x26200A  # Format32bppArgb

AVIF_HASINDEX = 0x10
AVIF_ISINTERLEAVED = 0x100
AVIIF_KEYFRAME = 0x10
SUPER_INDEX_ENTRIES = 1024          # room reserved in 'indx' (one entry per RIFF segment)
FIRST_SEGMENT_LIMIT = 1 << 30       # AVI 1.0 readers see the first RIFF only: keep it below 1 GiB
SEGMENT_LIMIT = 0x7F000000          # < 2 GiB per RIFF 'AVIX'


def _fourcc(s):
    return s.encode("ascii")


def get_rate_and_scale(frame_rate):
    """VideoStream.GetRateAndScale (VideoStream.cs:229-235)."""
    scale = 1
    rate = float(frame_rate)
    while rate != int(rate):
        rate = rate * 10
        scale *= 10
    return int(rate), scale


def _bits_per_pixel(fmt):
    if fmt in (PIXELFORMAT_32BPP_RGB, PIXELFORMAT_32BPP_ARGB):
        return 32
    if fmt == PIXELFORMAT_24BPP_RGB:
        return 24
    raise ValueError("unsupported PixelFormat 0x%x" % fmt)


def to_dib_bits(bmp, bpp):
    """Bitmap -> the bytes AddFrame hands to AVIStreamWrite: rows bottom-up, BGR(A), stride padded to 4."""
    a = np.asarray(bmp)
    if a.ndim == 2 and a.dtype == np.uint32:
        # packed R | G<<8 | B<<16 | A<<24 (little endian: bytes R, G, B, A) -> one byte gather per pixel
        h, w = a.shape
        nb = bpp // 8
        stride = (w * nb + 3) & ~3
        rgba = np.ascontiguousarray(a[::-1]).view(np.uint8).reshape(h, w, 4)
        if stride == w * nb:
            return rgba[..., [2, 1, 0, 3][:nb]].tobytes(), stride
        out = np.zeros((h, stride), dtype=np.uint8)
        out[:, : w * nb].reshape(h, w, nb)[...] = rgba[..., [2, 1, 0, 3][:nb]]
        return out.tobytes(), stride
    elif a.ndim == 3 and a.dtype == np.uint8 and a.shape[2] in (3, 4):
        r, g, b = a[..., 0], a[..., 1], a[..., 2]
        al = a[..., 3] if a.shape[2] == 4 else np.full(a.shape[:2], 255, np.uint8)
    else:
        raise ValueError("Bitmap must be (H, W) uint32 or (H, W, 3|4) uint8")
    h, w = a.shape[0], a.shape[1]
    nb = bpp // 8
    stride = (w * nb + 3) & ~3
    out = np.zeros((h, stride), dtype=np.uint8)
    px = out[:, : w * nb].reshape(h, w, nb)
    px[..., 0], px[..., 1], px[..., 2] = b, g, r
    if nb == 4:
        px[..., 3] = al
    return out[::-1].tobytes(), stride


class VideoStream:
    """aviFileWrapper_src/VideoStream.cs — the write side (and GetBitmap for files of this module)."""

    def __init__(self, manager, frame_rate, frame_size, width, height, fmt):
        self._m = manager
        self.FrameRate = float(frame_rate)
        self.Width, self.Height = int(width), int(height)
        self.CountBitsPerPixel = _bits_per_pixel(fmt)
        self._stride = (self.Width * (self.CountBitsPerPixel // 8) + 3) & ~3
        self.FrameSize = int(frame_size) if frame_size else self._stride * self.Height
        self.CountFrames = 0
        self.FirstFrame = 0
        self.WriteCompressed = False
        self._reader = None

    def AddFrame(self, bmp):
        """VideoStream.AddFrame (VideoStream.cs:344-365): flip, then write stride*height bytes."""
        if self._reader is not None:
            raise IOError("stream was opened for reading")
        a = np.asarray(bmp)
        if a.shape[0] != self.Height or a.shape[1] != self.Width:
            raise ValueError("frame size differs from the stream format")
        bits, stride = to_dib_bits(a, self.CountBitsPerPixel)
        assert stride == self._stride
        self._m._write_frame(bits)
        self.CountFrames += 1

    def AddFrameBits(self, bits):
        """The AVIStreamWrite call of AddFrame alone (VideoStream.cs:351-356): `bits` are the locked bitmap's
        bytes, i.e. rows bottom-up, BGR(A), Stride * Height of them.  For callers that convert on the device."""
        if self._reader is not None:
            raise IOError("stream was opened for reading")
        if len(bits) != self._stride * self.Height:
            raise ValueError("expected Stride * Height = %d bytes" % (self._stride * self.Height))
        self._m._write_frame(bits)
        self.CountFrames += 1

    def Flush(self):
        """Wait until every frame handed to AddFrame / AddFrameBits so far has reached the file (io_threads > 0)."""
        self._m._flush_payloads()

    def GetBitmap(self, position):
        """Frame `position` as an (H, W) uint32 array in the XNA Color layout (files of this module)."""
        if self._reader is None:
            raise IOError("stream was opened for writing")
        off, size = self._reader[position]
        with open(self._m.fileName, "rb") as f:
            f.seek(off)
            raw = np.frombuffer(f.read(size), dtype=np.uint8)
        nb = self.CountBitsPerPixel // 8
        rows = raw.reshape(self.Height, self._stride)[::-1, : self.Width * nb].reshape(self.Height, self.Width, nb).astype(np.uint32)
        alpha = rows[..., 3] if nb == 4 else np.uint32(255)
        return (rows[..., 2] | (rows[..., 1] << 8) | (rows[..., 0] << 16) | (alpha << 24)).astype(np.uint32)


_GROW_AHEAD = 256 << 20


def _write_payload(fd, bits, pos):
    """Worker of AviManager(io_threads > 0): copy one frame's bytes to offset `pos` of the file through a shared mapping."""
    import mmap

    n = len(bits)
    if n == 0:
        return None
    try:
        a = pos // mmap.ALLOCATIONGRANULARITY * mmap.ALLOCATIONGRANULARITY
        mm = mmap.mmap(fd, pos - a + n, offset=a)
    except (OSError, ValueError):
        return os.pwrite(fd, bits, pos)
    dst = None
    try:
        dst = np.frombuffer(mm, dtype=np.uint8, count=n, offset=pos - a)
        dst[:] = np.frombuffer(bits, dtype=np.uint8, count=n)   # numpy copies without the GIL
    finally:
        dst = None      # the mapping cannot be closed while a view of it exists
        mm.close()
    return n


class AviManager:
    """aviFileWrapper_src/AviManager.cs: `AviManager(fileName, open)`; open=False creates/overwrites."""

    def __init__(self, fileName, open_existing, first_segment_limit=FIRST_SEGMENT_LIMIT, segment_limit=SEGMENT_LIMIT, io_threads=0):
        """io_threads > 0: frame payloads are written by that many worker threads at the offset the frame was given (the
        chunk headers and indexes stay sequential) — appending 3.7 GB of frames is otherwise one memcpy stream.  The
        workers copy into a shared mapping of the file: write()/pwrite() calls on one file serialise on its inode lock
        (tmpfs on the bench box: 4.1-4.5 GB/s with 1, 4 or 8 writers), page faults of a mapping do not (7.0 GB/s with 8,
        tools/shmbench.py); a file system that cannot be mapped falls back to os.pwrite.
        The caller must keep a frame's buffer alive until VideoStream.Flush() / Close()."""
        self.fileName = fileName
        self._open = bool(open_existing)
        self._stream = None
        self._f = None
        self._pool, self._pending = None, []
        if io_threads > 0 and not open_existing:
            from concurrent.futures import ThreadPoolExecutor

            self._pool = ThreadPoolExecutor(max_workers=int(io_threads))
        self._limits = (int(first_segment_limit), int(segment_limit))
        if self._open:
            if not os.path.exists(fileName):
                raise IOError("Exception in AVIFileOpen: file not found")  # the wrapper throws on a non-zero HRESULT
            self._parse()
        else:
            self._f = open(fileName, "w+b")   # read access too: the payload workers map the file
            self._mapped_to = 0                # the file has been extended to this size for the mappings

    # ---- writing --------------------------------------------------------------------------------
    def AddVideoStream(self, isCompressed, frameRate, *args):
        """AddVideoStream(isCompressed, frameRate, frameSize, width, height, format)  (AviManager.cs:124)
        AddVideoStream(isCompressed, frameRate, firstFrame)                          (AviManager.cs:147)"""
        if self._open or self._stream is not None:
            raise IOError("Exception in AVIFileCreateStream")
        first = None
        if len(args) == 1:
            first = np.asarray(args[0])
            h, w = first.shape[0], first.shape[1]
            fmt = PIXELFORMAT_24BPP_RGB if (first.ndim == 2 or first.shape[2] == 3) else PIXELFORMAT_32BPP_ARGB
            frame_size = 0
        elif len(args) == 4:
            frame_size, w, h, fmt = args
        else:
            raise TypeError("AddVideoStream: unknown overload")
        self._stream = VideoStream(self, frameRate, frame_size, w, h, fmt)
        self._begin_file()
        if first is not None:
            self._stream.AddFrame(first)  # the Bitmap overload writes its first frame (VideoStream.cs:104-113)
        return self._stream

    def GetVideoStream(self):
        if self._stream is None:
            raise IOError("Exception in AVIFileGetStream")
        return self._stream

    def _begin_file(self):
        s, f = self._stream, self._f
        rate, scale = get_rate_and_scale(s.FrameRate)
        self._rate, self._scale = rate, scale
        f.write(_fourcc("RIFF") + struct.pack("<I", 0) + _fourcc("AVI "))
        hdrl_start = f.tell()
        f.write(_fourcc("LIST") + struct.pack("<I", 0) + _fourcc("hdrl"))
        self._avih_pos = f.tell()
        f.write(_fourcc("avih") + struct.pack("<I", 56) + b"\0" * 56)
        strl_start = f.tell()
        f.write(_fourcc("LIST") + struct.pack("<I", 0) + _fourcc("strl"))
        self._strh_pos = f.tell()
        f.write(_fourcc("strh") + struct.pack("<I", 56) + b"\0" * 56)
        # BITMAPINFOHEADER as SetFormat fills it (VideoStream.cs:373-384); biCompression = BI_RGB
        f.write(_fourcc("strf") + struct.pack("<I", 40) +
                struct.pack("<IiiHHIIiiII", 40, s.Width, s.Height, 1, s.CountBitsPerPixel, 0, s._stride * s.Height, 0, 0, 0, 0))
        self._indx_pos = f.tell()
        indx_size = 24 + 16 * SUPER_INDEX_ENTRIES
        f.write(_fourcc("indx") + struct.pack("<I", indx_size) + b"\0" * indx_size)
        end = f.tell()
        f.seek(strl_start + 4); f.write(struct.pack("<I", end - strl_start - 8)); f.seek(end)
        self._dmlh_pos = f.tell() + 20
        f.write(_fourcc("LIST") + struct.pack("<I", 4 + 8 + 248) + _fourcc("odml") + _fourcc("dmlh") + struct.pack("<I", 248) + b"\0" * 248)
        end = f.tell()
        f.seek(hdrl_start + 4); f.write(struct.pack("<I", end - hdrl_start - 8)); f.seek(end)
        self._segments = []          # (index chunk offset, index chunk size, frames)
        self._riff_start = 0
        self._idx1 = []              # legacy index of the first segment: (offset relative to 'movi', size)
        self._frames_first = 0
        self._begin_movi()

    def _begin_movi(self):
        f = self._f
        self._movi_start = f.tell()
        f.write(_fourcc("LIST") + struct.pack("<I", 0) + _fourcc("movi"))
        self._seg_entries = []       # (absolute offset of the chunk DATA, size)

    def _end_segment(self, last):
        f = self._f
        n = len(self._seg_entries)
        if n:
            base = self._movi_start
            ix_pos = f.tell()
            body = struct.pack("<HBBI4sQI", 2, 0, 1, n, _fourcc("00db"), base, 0)
            body += b"".join(struct.pack("<II", off - base, size) for off, size in self._seg_entries)
            f.write(_fourcc("ix00") + struct.pack("<I", len(body)) + body)
            self._segments.append((ix_pos, 8 + len(body), n))
        end = f.tell()
        f.seek(self._movi_start + 4); f.write(struct.pack("<I", end - self._movi_start - 8)); f.seek(end)
        if self._riff_start == 0:
            body = b"".join(struct.pack("<4sIII", _fourcc("00db"), AVIIF_KEYFRAME, off, size) for off, size in self._idx1)
            f.write(_fourcc("idx1") + struct.pack("<I", len(body)) + body)
            self._frames_first = len(self._idx1)
            end = f.tell()
        f.seek(self._riff_start + 4); f.write(struct.pack("<I", end - self._riff_start - 8)); f.seek(end)
        if not last:
            self._riff_start = f.tell()
            f.write(_fourcc("RIFF") + struct.pack("<I", 0) + _fourcc("AVIX"))
            self._begin_movi()

    def _write_frame(self, bits):
        f = self._f
        limit = self._limits[0] if self._riff_start == 0 else self._limits[1]
        if self._seg_entries and f.tell() - self._riff_start + len(bits) + 64 + 8 * (len(self._seg_entries) + 1) * 3 > limit:
            if len(self._segments) + 1 >= SUPER_INDEX_ENTRIES:
                raise IOError("Exception in VideoStreamWrite: too many RIFF segments")
            self._end_segment(last=False)
        pos = f.tell()
        f.write(_fourcc("00db") + struct.pack("<I", len(bits)))
        if self._pool is None:
            f.write(bits)
        else:
            # leave a hole of the payload's size and let a worker fill it
            f.seek(pos + 8 + len(bits))
            if pos + 8 + len(bits) > self._mapped_to:   # a mapping cannot reach past the end of the file: grow it ahead
                self._mapped_to = pos + 8 + len(bits) + _GROW_AHEAD
                os.ftruncate(f.fileno(), self._mapped_to)
            self._pending.append(self._pool.submit(_write_payload, f.fileno(), bits, pos + 8))
        if len(bits) & 1:
            f.write(b"\0")
        self._seg_entries.append((pos + 8, len(bits)))
        if self._riff_start == 0:
            self._idx1.append((pos - (self._movi_start + 8), len(bits)))

    def _flush_payloads(self):
        pending, self._pending = self._pending, []
        for fut in pending:
            n = fut.result()          # re-raises a failed write (disk full ...)
            if n is not None and n <= 0:
                raise IOError("Exception in VideoStreamWrite: short write")

    def Close(self):
        """AviManager.Close (AviManager.cs:327): finishes the headers and indexes."""
        if self._f is None:
            return
        self._flush_payloads()
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        if self._stream is not None:
            s, f = self._stream, self._f
            self._end_segment(last=True)
            self._end_of_file = f.tell()
            total = s.CountFrames
            usec = int(round(1e6 * self._scale / self._rate)) if self._rate else 0
            frame_bytes = s._stride * s.Height
            f.seek(self._avih_pos + 8)
            f.write(struct.pack("<IIIIIIIIII16x", usec, int(frame_bytes * self._rate / max(1, self._scale)), 0,
                                AVIF_HASINDEX | AVIF_ISINTERLEAVED, self._frames_first, 0, 1, frame_bytes, s.Width, s.Height))
            # AVISTREAMINFO as CreateStreamWithoutFormat fills it (VideoStream.cs:238-271)
            f.seek(self._strh_pos + 8)
            f.write(struct.pack("<4s4sIHHIIIIIIiIhhhh", _fourcc("vids"), _fourcc("DIB "), 0, 0, 0, 0, self._scale, self._rate, 0, total,
                                s.FrameSize, -1, 0, 0, 0, s.Width, s.Height))
            f.seek(self._indx_pos + 8)
            f.write(struct.pack("<HBBI4s12x", 4, 0, 0, len(self._segments), _fourcc("00db")))
            for off, size, dur in self._segments:
                f.write(struct.pack("<QII", off, size, dur))
            f.seek(self._dmlh_pos)
            f.write(struct.pack("<I", total))
            if self._mapped_to:
                f.flush()
                os.ftruncate(f.fileno(), self._end_of_file)   # drop what was grown ahead for the mappings
        self._f.close()
        self._f = None

    # ---- reading files written by this module ----------------------------------------------------
    def _parse(self):
        with open(self.fileName, "rb") as f:
            data_hdr = f.read(12)
            if data_hdr[:4] != b"RIFF" or data_hdr[8:12] != b"AVI ":
                raise IOError("Exception in AVIFileOpen: not an AVI file")
            f.seek(12 + 12)  # RIFF header + LIST hdrl header
            assert f.read(4) == b"avih"
            f.seek(4, 1)
            avih = struct.unpack("<IIIIIIIIII", f.read(40)); f.seek(16, 1)
            f.seek(12, 1)    # LIST strl
            assert f.read(4) == b"strh"
            f.seek(4, 1)
            strh = struct.unpack("<4s4sIHHIIIIIIiIhhhh", f.read(56))
            assert f.read(4) == b"strf"
            f.seek(4, 1)
            bih = struct.unpack("<IiiHHIIiiII", f.read(40))
            assert f.read(4) == b"indx"
            f.seek(4, 1)
            _, _, _, nseg, _ = struct.unpack("<HBBI4s12x", f.read(24))
            segs = [struct.unpack("<QII", f.read(16)) for _ in range(nseg)]
            frames = []
            for off, size, dur in segs:
                f.seek(off + 8)
                _, _, _, n, _, base, _ = struct.unpack("<HBBI4sQI", f.read(24))
                for _ in range(n):
                    o, sz = struct.unpack("<II", f.read(8))
                    frames.append((base + o, sz & 0x7FFFFFFF))
        self.MainHeader = dict(zip(("dwMicroSecPerFrame", "dwMaxBytesPerSec", "dwPaddingGranularity", "dwFlags", "dwTotalFrames",
                                    "dwInitialFrames", "dwStreams", "dwSuggestedBufferSize", "dwWidth", "dwHeight"), avih))
        self.StreamInfo = {"fccType": strh[0], "fccHandler": strh[1], "dwScale": strh[6], "dwRate": strh[7], "dwLength": strh[9],
                           "dwSuggestedBufferSize": strh[10], "dwQuality": strh[11], "rcFrame": strh[13:17]}
        fmt = PIXELFORMAT_24BPP_RGB if bih[4] == 24 else PIXELFORMAT_32BPP_ARGB
        s = VideoStream(self, strh[7] / max(1, strh[6]), bih[6], bih[1], bih[2], fmt)
        s.CountFrames = len(frames)
        s._reader = frames
        self._stream = s
        self.Segments = segs
