"""Decoding of compressed audio on the host, upstream of the hot path.

The reference opens a file with soundfile (libsndfile + mpg123) and falls back to PyAV = FFmpeg
(src/stream/audio.py:22-44).  Neither Python package is in this image, but an FFmpeg 8 build ships inside
`opencv_python_headless.libs/`; this module drives its libavformat / libavcodec directly through ctypes, i.e. the
reference's fallback decoder without PyAV.  Output: float32 [n] (mono) or [n, channels], at the file's own rate --
exactly what `track.read()` hands to `WorkerStreamer.queue_chunk` (src/stream/worker.py:110-129); downmix and
resampling then run on the GPU (`Engine.predict_pcm`).

Only four struct fields are read by offset, all stable since FFmpeg 5 (checked at import against the library's
major versions): AVFormatContext.streams (+48), AVStream.codecpar (+16), AVPacket.stream_index (+36),
AVFrame.{data[0..7] (+0), nb_samples (+112), format (+116)}.  Everything else goes through exported functions.
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
    if (avformat.avformat_version() >> 16) < 59 or (avcodec.avcodec_version() >> 16) < 59:
        raise RuntimeError("FFmpeg >= 5 required (struct offsets used by buzzdetect_b200.audio)")
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
    avutil.av_log_set_level.argtypes = [C.c_int]
    avutil.av_log_set_level(16)                                          # AV_LOG_ERROR: no per-file chatter
    _LIBS = (avutil, avcodec, avformat)
    return _LIBS


def decode_file(path: str) -> tuple[np.ndarray, int]:
    """Decode the first audio stream of `path` -> (float32 samples [n] or [n, channels], samplerate)."""
    avutil, avcodec, avformat = _load()
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    fmt = C.c_void_p()
    if avformat.avformat_open_input(C.byref(fmt), os.fsencode(path), None, None) < 0:
        raise ValueError(f"{path}: cannot open (unsupported container?)")
    ctx = C.c_void_p()
    pkt = C.c_void_p(avcodec.av_packet_alloc())
    frame = C.c_void_p(avutil.av_frame_alloc())
    try:
        if avformat.avformat_find_stream_info(fmt, None) < 0:
            raise ValueError(f"{path}: no stream information")
        dec = C.c_void_p()
        si = avformat.av_find_best_stream(fmt, _AVMEDIA_TYPE_AUDIO, -1, -1, C.byref(dec), 0)
        if si < 0 or not dec.value:
            raise ValueError(f"{path}: no decodable audio stream")
        streams = C.c_void_p.from_address(fmt.value + 48).value
        stream = C.c_void_p.from_address(streams + 8 * si).value
        codecpar = C.c_void_p.from_address(stream + 16).value
        ctx = C.c_void_p(avcodec.avcodec_alloc_context3(dec))
        if avcodec.avcodec_parameters_to_context(ctx, codecpar) < 0 or avcodec.avcodec_open2(ctx, dec, None) < 0:
            raise ValueError(f"{path}: cannot open the decoder")
        v = C.c_int64()
        if avutil.av_opt_get_int(ctx, b"ar", 0, C.byref(v)) < 0 or v.value <= 0:
            raise ValueError(f"{path}: unknown sample rate")
        rate = int(v.value)
        chunks = []

        def drain():
            while avcodec.avcodec_receive_frame(ctx, frame) == 0:
                n = C.c_int.from_address(frame.value + 112).value
                f = C.c_int.from_address(frame.value + 116).value
                if f not in _FMT or n <= 0:
                    continue
                dt, planar = _FMT[f]
                if planar:
                    planes = []
                    for ch in range(8):
                        p = C.c_void_p.from_address(frame.value + 8 * ch).value
                        if not p:
                            break
                        planes.append(np.frombuffer((C.c_char * (n * np.dtype(dt).itemsize)).from_address(p), dtype=dt).copy())
                    a = np.stack(planes, axis=1)
                else:
                    # packed: channel count from the linesize is not reliable; assume what the first plane holds
                    p = C.c_void_p.from_address(frame.value).value
                    ls = C.c_int.from_address(frame.value + 64).value
                    nch = max(1, ls // (n * np.dtype(dt).itemsize))
                    a = np.frombuffer((C.c_char * (n * nch * np.dtype(dt).itemsize)).from_address(p), dtype=dt).copy()
                    a = a.reshape(n, nch)
                chunks.append(_to_float(a))

        while avformat.av_read_frame(fmt, pkt) >= 0:
            if C.c_int.from_address(pkt.value + 36).value == si:
                avcodec.avcodec_send_packet(ctx, pkt)
                drain()
            avcodec.av_packet_unref(pkt)
        avcodec.avcodec_send_packet(ctx, None)
        drain()
    finally:
        if ctx:
            avcodec.avcodec_free_context(C.byref(ctx))
        avcodec.av_packet_free(C.byref(pkt))
        avutil.av_frame_free(C.byref(frame))
        avformat.avformat_close_input(C.byref(fmt))
    if not chunks:
        raise ValueError(f"{path}: decoded no audio")
    x = np.concatenate(chunks, axis=0)
    if x.shape[1] == 1:
        x = x[:, 0]
    return np.ascontiguousarray(x, dtype=np.float32), rate


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
