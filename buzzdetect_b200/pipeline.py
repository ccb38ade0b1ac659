"""File-level driver around the hot path: audio files in -> result CSVs out.

The narrow slice of the reference's streamer + inferer + writer that sits directly either side of the path, restated
without soundfile / librosa / pandas / TensorFlow:

  * WorkerStreamer._chunk_file / queue_chunk   src/stream/worker.py:61-135   chunk list, resume gaps, python-float
                                                                             sample indexing, bad reads end the file
  * WorkerInferer.process_chunk                src/inference/worker.py:71-74
  * WorkerWriter.write_results                 src/write/worker.py:67-87     append `<ident>_buzzpart.csv`; when the file
                                                                             is covered: sort by start, write
                                                                             `<ident>_buzzdetect.csv`, delete the partial

What changed against the reference's streamer (north star: "src/stream chunk batching now feeds pinned buffers"):
chunks are read straight into a ring of PINNED host buffers in their decoded form (int16 PCM for WAV, float32 for
FFmpeg-decoded formats) and go to the GPU as they are; downmix, resampling, frontend, CNN and head run on the device
(`Engine.submit_pcm`), several chunks in flight, small chunks coalesced into one pass.  Only [n_frames, 13] floats come
back.  Reader threads fill the ring ahead of the inferer (the reference runs 8-24 streamer threads for the same reason,
src/pipeline/coordination.py:129-135).
"""
from __future__ import annotations

import concurrent.futures
import csv
import logging
import os
import struct
import threading

import numpy as np

from . import capi, config as cfg, stream, weights as W, write

LOG = logging.getLogger("buzzdetect_b200")
BAD_READ_ALLOWANCE = 0.01       # src/config.py:17: tail fraction that may be corrupt before the message becomes a warning
FILE_SIZE_MINIMUM = 5000        # src/config.py:19: smaller files are skipped


# ----------------------------------------------------------------------------------------------- tracks
class WavTrack:
    """Minimal stand-in for the reference's AudioDriver (src/stream/driver.py:3-22) over RIFF/WAVE PCM16, with
    readinto() so a chunk lands in a pinned buffer without an intermediate bytes object."""

    fmt = 1                                     # int16 PCM for bd_submit_pcm_host
    dtype = np.int16

    def __init__(self, path: str):
        self._f = open(path, "rb", buffering=0)
        try:
            self._parse()
        except Exception:
            self._f.close()
            raise
        self._pos = 0

    def _parse(self):
        f = self._f
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError("not a RIFF/WAVE file")
        fmt_seen = False
        while True:
            hdr = f.read(8)
            if len(hdr) < 8:
                raise ValueError("WAVE file has no data chunk")
            cid, size = hdr[:4], struct.unpack("<I", hdr[4:])[0]
            if cid == b"fmt ":
                body = f.read(size + (size & 1))
                tag, ch, rate, _, _, bits = struct.unpack("<HHIIHH", body[:16])
                if tag == 0xFFFE and size >= 26:                # WAVE_FORMAT_EXTENSIBLE: the sub-format's first two bytes
                    tag = struct.unpack("<H", body[24:26])[0]
                if tag != 1 or bits != 16:
                    raise ValueError("only 16-bit PCM WAV is supported by this reader")
                self.samplerate, self.channels = int(rate), int(ch)
                fmt_seen = True
            elif cid == b"data":
                if not fmt_seen:
                    raise ValueError("WAVE data chunk before fmt chunk")
                self._data_off = f.tell()
                end = os.fstat(f.fileno()).st_size
                size = min(size, end - self._data_off) if size not in (0, 0xFFFFFFFF) else end - self._data_off
                self._frame_bytes = 2 * self.channels
                self.frames = size // self._frame_bytes
                return
            else:
                f.seek(size + (size & 1), os.SEEK_CUR)

    @property
    def duration(self) -> float:
        return self.frames / self.samplerate

    def seek(self, frame: int):
        self._pos = min(max(int(frame), 0), self.frames)

    def tell(self) -> int:
        return self._pos

    def readinto(self, out: np.ndarray, n: int) -> int:
        """Read up to n frames at the current position into `out` ([n] or [n, channels] int16, C-contiguous); returns
        the number of frames read.  Uses pread, so several reader threads may share one track."""
        n = max(0, min(int(n), self.frames - self._pos))
        mv = memoryview(out).cast("B")[: n * self._frame_bytes]
        got, off = 0, self._data_off + self._pos * self._frame_bytes
        while got < len(mv):
            k = os.preadv(self._f.fileno(), [mv[got:]], off + got)
            if k <= 0:
                break
            got += k
        frames = got // self._frame_bytes
        self._pos += frames
        return frames

    def read_at(self, out: np.ndarray, frame: int, n: int) -> int:
        """Positioned read (thread-safe: no shared cursor)."""
        frame = min(max(int(frame), 0), self.frames)
        n = max(0, min(int(n), self.frames - frame))
        mv = memoryview(out).cast("B")[: n * self._frame_bytes]
        got, off = 0, self._data_off + frame * self._frame_bytes
        while got < len(mv):
            k = os.preadv(self._f.fileno(), [mv[got:]], off + got)
            if k <= 0:
                break
            got += k
        return got // self._frame_bytes

    def read(self, n: int) -> np.ndarray:
        shape = (max(0, min(int(n), self.frames - self._pos)),) + ((self.channels,) if self.channels > 1 else ())
        out = np.empty(shape, dtype=np.int16)
        got = self.readinto(out, shape[0])
        return out[:got]

    def close(self):
        self._f.close()


class DecodedTrack:
    """Compressed formats (mp3, flac, ogg, m4a ...) and non-PCM16 WAVs: decoded incrementally on the host through FFmpeg
    (buzzdetect_b200.audio.StreamDecoder, the reference's PyAV fallback without PyAV, src/stream/audio.py:29-44) -- a
    chunk's worth of float32 at a time, duration from the container."""

    fmt = 0                                     # float32 PCM for bd_submit_pcm_host
    dtype = np.float32

    def __init__(self, path: str):
        from . import audio
        self._d = audio.StreamDecoder(path)
        self.samplerate = self._d.samplerate
        self.channels = max(1, self._d.channels)
        if self._d.frames is None:              # no duration in the container: count by decoding once
            n = 0
            while True:
                k = self._d.read(1 << 22).shape[0]
                if k == 0:
                    break
                n += k
            self.frames = n
            self._d.seek(0)
        else:
            self.frames = self._d.frames

    @property
    def duration(self) -> float:
        return self.frames / self.samplerate

    def seek(self, frame: int):
        self._d.seek(frame)

    def tell(self) -> int:
        return self._d.tell()

    def read(self, n: int) -> np.ndarray:
        a = self._d.read(n)
        self.channels = max(1, self._d.channels)
        return a

    def readinto(self, out: np.ndarray, n: int) -> int:
        a = self.read(n)
        out[: a.shape[0]] = a
        return a.shape[0]

    def close(self):
        self._d.close()


def open_track(path: str):
    """WAV PCM16 through the positioned reader, everything else through FFmpeg."""
    if path.lower().endswith((".wav", ".wave")):
        try:
            return WavTrack(path)
        except ValueError:
            pass
    return DecodedTrack(path)


# ----------------------------------------------------------------------------------------------- result files
def _fmt(v) -> str:
    """What pandas.to_csv prints for a float32/float64 cell: the shortest repr that round-trips in that dtype."""
    return str(v)


def _rows_text(header, start, values, with_header: bool) -> str:
    import io
    buf = io.StringIO()
    w = csv.writer(buf, lineterminator="\n")
    if with_header:
        w.writerow(header)
    vals = values if values.ndim == 2 else values[:, None]
    for s, row in zip(start, vals):
        w.writerow([_fmt(float(s))] + [_fmt(x) for x in row])
    return buf.getvalue()


def _append_rows(path: str, header: list[str], start: np.ndarray, values: np.ndarray):
    """write/worker.py:73-80: append a chunk's rows, header only when the file is new.  One O_APPEND write per chunk, and
    the header is written by whoever creates the file, so several ranks (one process per GPU) may share a partial file."""
    body = _rows_text(header, start, values, False).encode()
    try:
        fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_EXCL | os.O_APPEND, 0o644)
        body = _rows_text(header, start, values, True).encode()
    except FileExistsError:
        fd = os.open(path, os.O_WRONLY | os.O_APPEND)
    try:
        os.write(fd, body)
    finally:
        os.close(fd)


def _read_partial_starts(path: str) -> np.ndarray:
    with open(path, newline="") as f:
        r = csv.reader(f)
        rows = list(r)
    out = []
    for row in rows:
        if not row or row[0] == "start":        # the header (a second rank may have raced its own in: skip every copy)
            continue
        out.append(float(row[0]))
    return np.array(out, dtype=np.float64)


def _finalise(partial: str, complete: str):
    """write/worker.py:83-87: read the partial, sort by start, write the final file, remove the partial.  Guarded by an
    exclusive lock file: with chunk ranges of one file spread over ranks, the rank that sees the file covered last
    finalises it."""
    lock = partial + ".lock"
    try:
        fd = os.open(lock, os.O_WRONLY | os.O_CREAT | os.O_EXCL)
    except FileExistsError:
        return False
    try:
        if not os.path.exists(partial):
            return False
        with open(partial, newline="") as f:
            rows = [r for r in csv.reader(f) if r]
        header = rows[0]
        body = [r for r in rows[1:] if r[0] != "start"]
        body.sort(key=lambda r: float(r[0]))
        tmp = complete + ".tmp"
        with open(tmp, "w", newline="") as f:
            w = csv.writer(f, lineterminator="\n")
            w.writerow(header)
            w.writerows(body)
        os.replace(tmp, complete)
        os.remove(partial)
        return True
    finally:
        os.close(fd)
        os.remove(lock)


def validate_framehop(framehop_prop: float) -> int:
    """Patch hop in STFT frames.  The reference pads with int(framehop_s * 16000) samples but steps patches by
    int(round(100 * framehop_s)) frames (embedders/yamnet/features.py:66-71,99); the two only agree -- and time stamps
    only stay exact -- for hops that are a whole number of 10 ms frames.  Others are rejected instead of drifting."""
    hop = capi.hop_frames_for(framehop_prop)
    framehop_s = 0.96 * framehop_prop
    if hop < 1 or hop > 96 or abs(hop * 0.01 - framehop_s) > 1e-9 or int(framehop_s * 16000) != hop * 160:
        raise ValueError(f"framehop_prop={framehop_prop} is not a whole number of 10 ms STFT frames")
    return hop


class _PinnedRing:
    """A few pinned chunk buffers per (engine, dtype, size), allocated once and re-used for every chunk and file."""

    def __init__(self):
        self._bufs = {}
        self._lock = threading.Lock()

    def get(self, key, index: int, shape, dtype):
        with self._lock:
            k = (key, index, np.dtype(dtype).str)
            b = self._bufs.get(k)
            need = int(np.prod(shape))
            if b is None or b.size < need:
                b = capi.pinned_empty(need, dtype)
                self._bufs[k] = b
            return b[:need].reshape(shape)


_RING = _PinnedRing()


def _handle_bad_read(ident: str, final_second: float, duration: float, log):
    """src/stream/worker.py:41-59."""
    msg = f"Unreadable audio at {round(final_second, 1)}s out of {round(duration, 1)}s for {ident}."
    if duration > 0 and 1 - (final_second / duration) > BAD_READ_ALLOWANCE:
        log(msg + "\nAborting early due to corrupt audio data.", "WARNING")
    else:
        log(msg + "\nBad audio is near file end, results should be mostly unaffected.", "DEBUG")


def _default_log(msg: str, level: str):
    LOG.log({"DEBUG": logging.DEBUG, "INFO": logging.INFO, "PROGRESS": logging.INFO, "WARNING": logging.WARNING,
             "ERROR": logging.ERROR}.get(level, logging.INFO), msg)


def analyze_wav(path_audio: str, dir_out: str, engine: "capi.Engine", classes: list[str], chunklength: float = 199.68,
                framehop_prop: float = 1.0, threshold: float | None = None, classes_keep="all",
                digits_results: int = 2, n_in_flight: int = 8, only_chunks=None, readers: int = 2, log=None,
                stop_event: threading.Event | None = None) -> dict:
    """One audio file (any format) through the path.  Resumes from `<ident>_buzzpart.csv` if present; skips finished
    files.  only_chunks: the (start, end) chunks of the file that THIS rank processes (chunk-range sharding of one long
    file over several GPUs, taken from make_plan(...).chunks -- the chunks themselves, not indices: the chunk list
    re-derived here shrinks as other ranks append their rows); the file is finalised by whichever rank finds it covered.
    Returns a small report."""
    log = log or _default_log
    ident = os.path.splitext(os.path.basename(path_audio))[0]
    partial = os.path.join(dir_out, ident + cfg.SUFFIX_RESULT_PARTIAL)
    complete = os.path.join(dir_out, ident + cfg.SUFFIX_RESULT_COMPLETE)
    os.makedirs(dir_out, exist_ok=True)
    report = {"ident": ident, "chunks": 0, "frames": 0, "skipped": False, "bad_read": False}
    if os.path.exists(complete):
        log(f"Skipping {ident}; already analyzed", "DEBUG")
        return dict(report, skipped=True)
    if os.path.getsize(path_audio) < FILE_SIZE_MINIMUM:
        log(f"Skipping {ident}; below minimum analyzeable size", "DEBUG")
        return dict(report, skipped=True)
    prov = getattr(engine, "weights_provenance", "caller")
    if prov.startswith("synthetic") and not W.synthetic_allowed():
        raise RuntimeError("refusing to write detections from SYNTHETIC YAMNet weights: provide the real blob "
                           "(BUZZ_YAMNET_WEIGHTS) or opt in with BUZZ_B200_ALLOW_SYNTHETIC=1 (tests / benchmarks)")
    framelength_s = 0.96
    framehop_s = framelength_s * framehop_prop
    hop_frames = validate_framehop(framehop_prop)
    chunklength = stream.setup_chunklength(chunklength, framelength_s)
    track = open_track(path_audio)
    covered = _read_partial_starts(partial) if os.path.exists(partial) else None
    chunklist = stream.file_chunklist(track.duration, chunklength, covered, framelength_s)
    if covered is not None and covered.size and not chunklist:
        log(f"Discovered non-cleaned file at {ident}; cleaning results", "DEBUG")
        _finalise(partial, complete)
        track.close()
        return report
    if only_chunks is not None:
        chunklist = sorted((float(a), float(b)) for a, b in only_chunks)
    mine = list(range(len(chunklist)))
    sr, ch = track.samplerate, track.channels
    n_ring = max(2, min(int(n_in_flight), engine.n_slots))
    positioned = isinstance(track, WavTrack)     # pread: reader threads work ahead; a decoder is a sequential stream
    pool = concurrent.futures.ThreadPoolExecutor(max_workers=max(1, readers)) if positioned and readers > 0 else None
    pending = []                                 # (ticket, chunk) in submission order
    frames_total = 0

    def load(k: int, j: int):
        """chunk j of the list -> ring buffer k; (buffer view, frames read, frames wanted)."""
        sample_from, read_size = stream.chunk_sample_range(chunklist[j], sr)
        buf = _RING.get(id(engine), k, (read_size, ch) if ch > 1 else (read_size,), track.dtype)
        if positioned:
            got = track.read_at(buf, sample_from, read_size)
        else:
            track.seek(sample_from)
            got = track.readinto(buf, read_size)
        return buf, got, read_size

    def drain_one():
        nonlocal frames_total
        tk, chunk = pending.pop(0)
        act = tk.result()
        frames_total += act.shape[0]
        if threshold is None:
            cols, start, vals = write.format_activations(act, classes, framehop_s, 2, time_start=chunk[0],
                                                         classes_keep=classes_keep, digits_results=digits_results)
        else:
            cols, start, vals = write.format_detections(act, threshold, classes, framehop_s, 2, chunk[0])
        _append_rows(partial, cols, start, vals)

    try:
        ahead = {}                               # position in `mine` -> future / result of its load
        depth = n_ring - 1 if pool else 0        # buffers being filled ahead of the one in use

        def schedule(pos):
            if pool and pos < len(mine) and pos not in ahead:
                ahead[pos] = pool.submit(load, pos % n_ring, mine[pos])

        for pos in range(min(depth, len(mine))):
            schedule(pos)
        for pos, j in enumerate(mine):
            if stop_event is not None and stop_event.is_set():
                break                            # src/stream/worker.py:145-146: bail between chunks
            # the ring buffer about to be (re)filled must have left the host: its ticket is n_ring submissions old
            while len(pending) >= n_ring - depth:
                drain_one()
            if pool:
                buf, n_read, read_size = ahead.pop(pos).result()
            else:
                buf, n_read, read_size = load(pos % n_ring, j)
            chunk = chunklist[j]
            short = n_read < read_size
            if short:
                # bad read: truncate the chunk, report, stop the file (src/stream/worker.py:119-126)
                _handle_bad_read(ident, (stream.chunk_sample_range(chunk, sr)[0] + n_read) / sr, track.duration, log)
                chunk = (chunk[0], round(chunk[0] + (n_read / sr), 1))
                report["bad_read"] = True
            if n_read > 0:
                tk = engine.submit_pcm(buf[:n_read], sr, hop_frames)
                pending.append((tk, chunk))
                report["chunks"] += 1
            if short:
                break
            schedule(pos + depth)
        while pending:
            drain_one()
    finally:
        for tk, _ in pending:                    # an exception above must not leave slots in flight
            try:
                tk.wait()
            except Exception:
                pass
        if pool:
            pool.shutdown(wait=True, cancel_futures=True)
        track.close()
    report["frames"] = frames_total
    if os.path.exists(partial) and not report["bad_read"] and not (stop_event is not None and stop_event.is_set()):
        # finished when nothing is left to do for the whole file (other ranks may still be working on their ranges)
        left = stream.file_chunklist(track.duration, chunklength, _read_partial_starts(partial), framelength_s)
        if not left:
            _finalise(partial, complete)
    elif os.path.exists(partial) and report["bad_read"] and only_chunks is None:
        _finalise(partial, complete)             # the reference marks the truncated chunk as the file's last chunk
    return report


analyze_file = analyze_wav


def make_plan(paths: list[str], dir_out: str, world_size: int, chunklength: float = 199.68,
              framehop_prop: float = 1.0) -> list[list[tuple[int, int]]]:
    """Work list per rank: [(file index, chunk index)], from the files' durations and whatever partial results exist.
    Files are dealt whole while there are at least as many as ranks, otherwise chunk RANGES of each file are spread over
    the ranks (buzzdetect_b200.shard).  Compute it ONCE (rank 0) and hand the same plan to every rank: a rank that
    re-derived it later would see rows the others have already appended."""
    from . import shard
    chunklength = stream.setup_chunklength(chunklength)
    lists = []
    for p in paths:
        ident = os.path.splitext(os.path.basename(p))[0]
        if os.path.exists(os.path.join(dir_out, ident + cfg.SUFFIX_RESULT_COMPLETE)) or os.path.getsize(p) < FILE_SIZE_MINIMUM:
            lists.append([])
            continue
        t = open_track(p)
        try:
            partial = os.path.join(dir_out, ident + cfg.SUFFIX_RESULT_PARTIAL)
            covered = _read_partial_starts(partial) if os.path.exists(partial) else None
            lists.append(stream.file_chunklist(t.duration, chunklength, covered))
        finally:
            t.close()
    plan = _Plan([(w.file_index, w.chunk_index) for w in r] for r in shard.plan(lists, world_size))
    plan.chunks = {(fi, ci): c for fi, l in enumerate(lists) for ci, c in enumerate(l)}
    return plan


class _Plan(list):
    """[rank] -> [(file index, chunk index)]; .chunks maps (file index, chunk index) -> (start, end) seconds."""
    chunks: dict


def analyze_files(paths: list[str], dir_out: str, rank: int = 0, world_size: int = 1, device: int | None = None,
                  plan=None, model=None, **kw) -> list[dict]:
    """This rank's share of `paths` on its GPU (one process per GPU, no collective: SURVEY.md section 8e)."""
    from .inference.models import load_model
    if plan is None:
        plan = make_plan(paths, dir_out, world_size, kw.get("chunklength", 199.68), kw.get("framehop_prop", 1.0))
    work = {}
    for fi, ci in plan[rank]:
        work.setdefault(fi, []).append(ci)
    if device is not None:
        os.environ["BUZZ_B200_DEVICE"] = str(device)
    if model is None:
        model = load_model(cfg.DEFAULT_MODEL, framehop_prop=kw.get("framehop_prop", 1.0), initialize=True)
    whole = {fi for fi in work if all(fi not in {f for f, _ in plan[r]} for r in range(world_size) if r != rank)}
    out = []
    for fi in sorted(work):
        out.append(analyze_wav(paths[fi], dir_out, model.model, model.config["classes"],
                               only_chunks=None if fi in whole else [plan.chunks[(fi, ci)] for ci in work[fi]], **kw))
    return out
