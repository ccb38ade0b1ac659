"""Decoding of compressed audio on the host, upstream of the hot path.

The reference opens a file with soundfile (libsndfile + mpg123) and falls back to PyAV = FFmpeg
(src/stream/audio.py:22-44).  Neither Python package is in this image, but an FFmpeg 8 build ships inside
`opencv_python_headless.libs/`; this module drives its libavformat / libavcodec directly through ctypes, i.e. the
reference's fallback decoder without PyAV.  Output: float32 [n] (mono) or [n, channels], at the file's own rate --
exactly what `track.read()` hands to `WorkerStreamer.queue_chunk` (src/stream/worker.py:110-129); downmix and
resampling then run on the GPU (`Engine.predict_pcm`).

A handful of struct fields are read by offset, stable from FFmpeg 6 (libavformat 60, checked at import):
AVFormatContext.streams (+48), AVStream.{codecpar (+16), time_base (+32), duration (+48)}, AVPacket.stream_index (+36),
AVFrame.{data[0..7] (+0), linesize[0] (+64), extended_data (+96), nb_samples (+112), format (+116)}.  The channel count
comes from the codec context's "ch_layout" option, the sample rate from "ar".  Everything else goes through exported
functions.  Decoding is incremental (StreamDecoder): a chunk's worth of samples at a time, duration from the container.
Gapless metadata (LAME start/end padding) is honoured by libavcodec itself, as it is for PyAV.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

_LIBS = None
_AVMEDIA_TYPE_AUDIO = 1
# AVSampleFormat -> (numpy dtype, planar, scale)
_FMT = {0: (np.uint8, False), 1: (np.int16, False), 2: (np.int32, False), 3: (np.float32, False), 4: (np.float64, False),
        5: (np.uint8, True), 6: (np.int16, True), 7: (np.int32, True), 8: (np.float32, True), 9: (np.float64, True)}


def _find_lib_dir() -> str:
    env = os.environ.get("BUZZ_FFMPEG_LIBDIR")
    if env:
        return env
    try:
        import cv2                                                     # noqa: F401  (only to locate its bundled libs)
        d = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
        if os.path.isdir(d):
            return d
    except ImportError:
        pass
    raise RuntimeError("no FFmpeg libraries found: set BUZZ_FFMPEG_LIBDIR to a directory holding libavformat/"
                       "libavcodec/libavutil (FFmpeg >= 5)")


def _load():
    global _LIBS
    if _LIBS is not None:
        return _LIBS
    d = _find_lib_dir()

    def load(stem, required=True):
        hits = sorted(glob.glob(os.path.join(d, stem + "*.so*")))
        if not hits:
            if required:
                raise RuntimeError(f"{stem} not found in {d}")
            return None
        return C.CDLL(hits[0], mode=C.RTLD_GLOBAL)

    for dep in ("libdrm", "libcrypto", "libssl", "libvpx", "libaom", "libpng16"):     # bundled deps of the codecs
        try:
            load(dep, required=False)
        except OSError:
            pass
    avutil = load("libavutil")
    load("libswresample", required=False)
    avcodec = load("libavcodec")
    avformat = load("libavformat")
    if (avformat.avformat_version() >> 16) < 60 or (avcodec.avcodec_version() >> 16) < 60:
        # AVStream gained its leading av_class pointer in libavformat 60: codecpar sits at +16 only from there on
        raise RuntimeError("FFmpeg >= 6 (libavformat >= 60) required: struct offsets used by buzzdetect_b200.audio")
    avformat.avformat_open_input.argtypes = [C.POINTER(C.c_void_p), C.c_char_p, C.c_void_p, C.c_void_p]
    avformat.avformat_find_stream_info.argtypes = [C.c_void_p, C.c_void_p]
    avformat.av_find_best_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_int]
    avformat.av_read_frame.argtypes = [C.c_void_p, C.c_void_p]
    avformat.avformat_close_input.argtypes = [C.POINTER(C.c_void_p)]
    avcodec.avcodec_alloc_context3.restype = C.c_void_p
    avcodec.avcodec_alloc_context3.argtypes = [C.c_void_p]
    avcodec.avcodec_parameters_to_context.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.avcodec_open2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    avcodec.avcodec_free_context.argtypes = [C.POINTER(C.c_void_p)]
    avcodec.av_packet_alloc.restype = C.c_void_p
    avcodec.av_packet_unref.argtypes = [C.c_void_p]
    avcodec.av_packet_free.argtypes = [C.POINTER(C.c_void_p)]
    avcodec.avcodec_send_packet.argtypes = [C.c_void_p, C.c_void_p]
    avcodec.avcodec_receive_frame.argtypes = [C.c_void_p, C.c_void_p]
    avutil.av_frame_alloc.restype = C.c_void_p
    avutil.av_frame_free.argtypes = [C.POINTER(C.c_void_p)]
    avutil.av_opt_get_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]
    avutil.av_opt_get_chlayout.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
    avutil.av_channel_layout_uninit.argtypes = [C.c_void_p]
    avutil.av_channel_layout_uninit.restype = None
    avutil.av_log_set_level.argtypes = [C.c_int]
    avutil.av_log_set_level(16)                                          # AV_LOG_ERROR: no per-file chatter
    _LIBS = (avutil, avcodec, avformat)
    return _LIBS


class StreamDecoder:
    """Incremental decoder over the first audio stream of a file: the demuxer and the decoder stay open and a chunk's
    worth of samples is produced at a time (a day-long mp3 is never held in memory).  Mirrors what the reference's
    AudioDriver offers to WorkerStreamer (src/stream/driver.py:3-22): samplerate, channels, frames, seek(), read().

    A decode error or a truncated file ends the stream early (read() returns fewer samples than asked), which is exactly
    the reference's "bad read" condition (src/stream/worker.py:119-126)."""

    def __init__(self, path: str):
        avutil, avcodec, avformat = _load()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self._libs = (avutil, avcodec, avformat)
        self._fmt = C.c_void_p()
        self._ctx = C.c_void_p()
        self._pkt = C.c_void_p()
        self._frame = C.c_void_p()
        self._open()

    # ------------------------------------------------------------------ open / close
    def _open(self):
        avutil, avcodec, avformat = self._libs
        path = self.path
        if avformat.avformat_open_input(C.byref(self._fmt), os.fsencode(path), None, None) < 0:
            self._fmt = C.c_void_p()
            raise ValueError(f"{path}: cannot open (unsupported container?)")
        try:
            self._pkt = C.c_void_p(avcodec.av_packet_alloc())
            self._frame = C.c_void_p(avutil.av_frame_alloc())
            if avformat.avformat_find_stream_info(self._fmt, None) < 0:
                raise ValueError(f"{path}: no stream information")
            dec = C.c_void_p()
            si = avformat.av_find_best_stream(self._fmt, _AVMEDIA_TYPE_AUDIO, -1, -1, C.byref(dec), 0)
            if si < 0 or not dec.value:
                raise ValueError(f"{path}: no decodable audio stream")
            self._si = si
            streams = C.c_void_p.from_address(self._fmt.value + 48).value
            stream = C.c_void_p.from_address(streams + 8 * si).value
            codecpar = C.c_void_p.from_address(stream + 16).value
            self._ctx = C.c_void_p(avcodec.avcodec_alloc_context3(dec))
            if avcodec.avcodec_parameters_to_context(self._ctx, codecpar) < 0 or avcodec.avcodec_open2(self._ctx, dec, None) < 0:
                raise ValueError(f"{path}: cannot open the decoder")
            v = C.c_int64()
            if avutil.av_opt_get_int(self._ctx, b"ar", 0, C.byref(v)) < 0 or v.value <= 0:
                raise ValueError(f"{path}: unknown sample rate")
            self.samplerate = int(v.value)
            # channel count from the codec context (the per-frame linesize is padded and cannot be trusted)
            lay = (C.c_char * 24)()
            self.channels = 0
            if avutil.av_opt_get_chlayout(self._ctx, b"ch_layout", 0, lay) >= 0:
                self.channels = int(C.c_int.from_buffer(lay, 4).value)
                avutil.av_channel_layout_uninit(lay)
            # container metadata: AVStream.time_base (+32), duration (+48) -- no decode needed for the chunk list
            tb_num = C.c_int.from_address(stream + 32).value
            tb_den = C.c_int.from_address(stream + 36).value
            dur = C.c_int64.from_address(stream + 48).value
            self.frames = None
            if dur > 0 and tb_num > 0 and tb_den > 0:
                self.frames = int(round(dur * tb_num / tb_den * self.samplerate))
        except Exception:
            self.close()
            raise
        self._buf = []            # decoded blocks not yet handed out
        self._buffered = 0
        self._pos = 0             # index of the next sample read() returns
        self._eof = False
        self.decode_errors = 0

    def close(self):
        avutil, avcodec, avformat = self._libs
        if self._ctx:
            avcodec.avcodec_free_context(C.byref(self._ctx))
            self._ctx = C.c_void_p()
        if self._pkt:
            avcodec.av_packet_free(C.byref(self._pkt))
            self._pkt = C.c_void_p()
        if self._frame:
            avutil.av_frame_free(C.byref(self._frame))
            self._frame = C.c_void_p()
        if self._fmt:
            avformat.avformat_close_input(C.byref(self._fmt))
            self._fmt = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ decoding
    def _take_frames(self) -> bool:
        """Collect every frame the decoder has ready; False on a decode error."""
        avutil, avcodec, avformat = self._libs
        frame = self._frame
        while True:
            rc = avcodec.avcodec_receive_frame(self._ctx, frame)
            if rc == _EAGAIN or rc == _EOF:
                return True
            if rc < 0:
                return False
            n = C.c_int.from_address(frame.value + 112).value
            f = C.c_int.from_address(frame.value + 116).value
            if f not in _FMT or n <= 0:
                continue
            dt, planar = _FMT[f]
            isz = np.dtype(dt).itemsize
            if planar:
                nch = self.channels
                if nch <= 0:                                       # no layout option: count the non-null planes
                    nch = sum(1 for ch in range(8) if C.c_void_p.from_address(frame.value + 8 * ch).value)
                    self.channels = nch
                ext = C.c_void_p.from_address(frame.value + 96).value      # extended_data: all planes, also beyond 8
                planes = []
                for ch in range(nch):
                    p = C.c_void_p.from_address(ext + 8 * ch).value
                    planes.append(np.frombuffer((C.c_char * (n * isz)).from_address(p), dtype=dt).copy())
                a = np.stack(planes, axis=1)
            else:
                p = C.c_void_p.from_address(frame.value).value
                if self.channels <= 0:                             # fall back to the first frame's (padded) linesize
                    ls = C.c_int.from_address(frame.value + 64).value
                    self.channels = max(1, ls // (n * isz))
                nch = self.channels
                a = np.frombuffer((C.c_char * (n * nch * isz)).from_address(p), dtype=dt).copy().reshape(n, nch)
            self._buf.append(_to_float(a))
            self._buffered += n

    def _pump(self) -> bool:
        """Decode one more packet into the buffer; False when the stream has ended (or broke)."""
        avutil, avcodec, avformat = self._libs
        if self._eof:
            return False
        while True:
            if avformat.av_read_frame(self._fmt, self._pkt) < 0:
                avcodec.avcodec_send_packet(self._ctx, None)       # flush
                self._take_frames()
                self._eof = True
                return False
            mine = C.c_int.from_address(self._pkt.value + 36).value == self._si
            ok = True
            if mine:
                rc = avcodec.avcodec_send_packet(self._ctx, self._pkt)
                ok = rc >= 0 or rc == _EAGAIN
                ok = self._take_frames() and ok
            avcodec.av_packet_unref(self._pkt)
            if not ok:
                # corrupt audio: the reference stops the file at a bad read (src/stream/worker.py:119-126)
                self.decode_errors += 1
                self._eof = True
                return False
            if mine:
                return True

    def read(self, n: int) -> np.ndarray:
        """Up to n frames from the current position: float32 [m] (mono) or [m, channels]; m < n at the end of the
        stream or at corrupt data."""
        while self._buffered < n and self._pump():
            pass
        if not self._buf:
            return np.zeros((0,) if self.channels <= 1 else (0, self.channels), dtype=np.float32)
        x = np.concatenate(self._buf, axis=0) if len(self._buf) > 1 else self._buf[0]
        out, rest = x[:n], x[n:]
        self._buf = [rest] if rest.shape[0] else []
        self._buffered = rest.shape[0]
        self._pos += out.shape[0]
        if out.ndim == 2 and out.shape[1] == 1:
            out = out[:, 0]
        return np.ascontiguousarray(out, dtype=np.float32)

    def tell(self) -> int:
        return self._pos

    def seek(self, frame: int):
        """Sample-accurate seek by decoding forward and discarding (chunk lists are ascending, so this is sequential in
        practice); a backward seek reopens the file."""
        frame = max(0, int(frame))
        if frame < self._pos:
            self.close()
            self._open()
        while self._pos < frame:
            got = self.read(min(frame - self._pos, 1 << 20)).shape[0]
            if got == 0:
                break


_EAGAIN = -11                    # AVERROR(EAGAIN) on Linux
_EOF = -541478725                # AVERROR_EOF = FFERRTAG('E','O','F',' ')


def decode_file(path: str) -> tuple[np.ndarray, int]:
    """Decode the first audio stream of `path` -> (float32 samples [n] or [n, channels], samplerate)."""
    d = StreamDecoder(path)
    try:
        parts = []
        while True:
            a = d.read(1 << 22)
            if a.shape[0] == 0:
                break
            parts.append(a)
        rate = d.samplerate
    finally:
        d.close()
    if not parts:
        raise ValueError(f"{path}: decoded no audio")
    return np.ascontiguousarray(np.concatenate(parts, axis=0), dtype=np.float32), rate


def _to_float(a: np.ndarray) -> np.ndarray:
    if a.dtype == np.float32:
        return a
    if a.dtype == np.float64:
        return a.astype(np.float32)
    if a.dtype == np.int16:
        return a.astype(np.float32) / 32768.0
    if a.dtype == np.int32:
        return (a.astype(np.float64) / 2147483648.0).astype(np.float32)
    if a.dtype == np.uint8:
        return (a.astype(np.float32) - 128.0) / 128.0
    raise ValueError(a.dtype)
