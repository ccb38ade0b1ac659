"""File-level driver around the hot path: WAV in -> result CSVs out.

This is the narrow slice of the reference's streamer + inferer + writer that sits directly either side of the path,
restated without soundfile / librosa / pandas / TensorFlow:

  * WorkerStreamer._chunk_file / queue_chunk   src/stream/worker.py:61-135   (chunk list, resume gaps, float->int
    sample indexing, short reads end the file)
  * WorkerInferer.process_chunk                src/inference/worker.py:71-74
  * WorkerWriter.write_results                 src/write/worker.py:67-87     (append `<ident>_buzzpart.csv`; when the
    file is complete: sort by start, write `<ident>_buzzdetect.csv`, delete the partial)

Audio is read with the stdlib `wave` module (PCM16 WAV) or, for compressed formats such as the reference's own
`audio_in/testbuzz.mp3`, decoded through FFmpeg via ctypes (buzzdetect_b200/audio.py), and handed to the GPU in its decoded form: downmix + resample + frontend + CNN + head run in one C-ABI call per chunk
(`bd_submit_pcm_host`), several chunks in flight.  The coordinator / logging / manifest layers are NOT rebuilt here.
"""
from __future__ import annotations

import csv
import os
import wave

import numpy as np

from . import capi, config as cfg, stream, write


class WavTrack:
    """Minimal stand-in for the reference's AudioDriver (src/stream/driver.py:3-22) over RIFF/WAVE PCM16."""

    fmt = 1                                     # int16 PCM for bd_submit_pcm_host

    def __init__(self, path: str):
        self._w = wave.open(path, "rb")
        if self._w.getsampwidth() != 2 or self._w.getcomptype() != "NONE":
            raise ValueError(f"{path}: only 16-bit PCM WAV is supported by this reader")
        self.samplerate = self._w.getframerate()
        self.channels = self._w.getnchannels()
        self.frames = self._w.getnframes()

    @property
    def duration(self) -> float:
        return self.frames / self.samplerate

    def seek(self, frame: int):
        self._w.setpos(min(max(frame, 0), self.frames))

    def read(self, n: int) -> np.ndarray:
        raw = self._w.readframes(n)
        a = np.frombuffer(raw, dtype="<i2")
        return a.reshape(-1, self.channels) if self.channels > 1 else a

    def close(self):
        self._w.close()


class DecodedTrack:
    """Compressed formats (mp3, flac, ogg, m4a ...): decoded once on the host through FFmpeg (buzzdetect_b200.audio,
    the reference's PyAV fallback without PyAV, src/stream/audio.py:29-44) and served from memory as float32."""

    fmt = 0                                     # float32 PCM for bd_submit_pcm_host

    def __init__(self, path: str):
        from . import audio
        self._x, self.samplerate = audio.decode_file(path)
        self.channels = 1 if self._x.ndim == 1 else self._x.shape[1]
        self.frames = self._x.shape[0]
        self._pos = 0

    @property
    def duration(self) -> float:
        return self.frames / self.samplerate

    def seek(self, frame: int):
        self._pos = min(max(frame, 0), self.frames)

    def read(self, n: int) -> np.ndarray:
        a = self._x[self._pos:self._pos + n]
        self._pos += a.shape[0]
        return a

    def close(self):
        self._x = None


def open_track(path: str):
    """WAV PCM16 through the stdlib reader, everything else through FFmpeg."""
    if path.lower().endswith(".wav"):
        try:
            return WavTrack(path)
        except (ValueError, wave.Error):
            pass
    return DecodedTrack(path)


def _fmt(v) -> str:
    """What pandas.to_csv prints for a float32/float64 cell: the shortest repr that round-trips in that dtype."""
    return str(v)


def _append_rows(path: str, header: list[str], start: np.ndarray, values: np.ndarray):
    new = not os.path.exists(path)
    with open(path, "a", newline="") as f:
        w = csv.writer(f, lineterminator="\n")
        if new:
            w.writerow(header)
        vals = values if values.ndim == 2 else values[:, None]
        for s, row in zip(start, vals):
            w.writerow([_fmt(float(s))] + [_fmt(x) for x in row])


def _read_partial_starts(path: str) -> np.ndarray:
    with open(path, newline="") as f:
        r = csv.DictReader(f)
        return np.array([float(row["start"]) for row in r], dtype=np.float64)


def _finalise(partial: str, complete: str):
    """write/worker.py:83-87: read the partial, sort by start, write the final file, remove the partial."""
    with open(partial, newline="") as f:
        rows = list(csv.reader(f))
    header, body = rows[0], rows[1:]
    body.sort(key=lambda r: float(r[0]))
    with open(complete, "w", newline="") as f:
        w = csv.writer(f, lineterminator="\n")
        w.writerow(header)
        w.writerows(body)
    os.remove(partial)


def analyze_wav(path_audio: str, dir_out: str, engine: "capi.Engine", classes: list[str], chunklength: float = 1198.08,
                framehop_prop: float = 1.0, threshold: float | None = None, classes_keep="all",
                digits_results: int = 2, n_in_flight: int = 2) -> dict:
    """One file through the path.  Resumes from `<ident>_buzzpart.csv` if present; skips finished files."""
    ident = os.path.splitext(os.path.basename(path_audio))[0]
    partial = os.path.join(dir_out, ident + cfg.SUFFIX_RESULT_PARTIAL)
    complete = os.path.join(dir_out, ident + cfg.SUFFIX_RESULT_COMPLETE)
    os.makedirs(dir_out, exist_ok=True)
    if os.path.exists(complete):
        return {"ident": ident, "chunks": 0, "frames": 0, "skipped": True}
    framelength_s = 0.96
    framehop_s = framelength_s * framehop_prop
    hop_frames = capi.hop_frames_for(framehop_prop)
    chunklength = stream.setup_chunklength(chunklength, framelength_s)
    track = open_track(path_audio)
    covered = _read_partial_starts(partial) if os.path.exists(partial) else None
    chunklist = stream.file_chunklist(track.duration, chunklength, covered, framelength_s)
    if covered is not None and not chunklist:
        _finalise(partial, complete)
        track.close()
        return {"ident": ident, "chunks": 0, "frames": 0, "skipped": False}

    n_slots = max(1, min(n_in_flight, engine.n_slots))
    pending = []            # (slot, chunk, act buffer, pcm keep-alive)
    frames_total = 0

    def drain_one():
        nonlocal frames_total
        slot, chunk, act, _keep = pending.pop(0)
        engine.wait(slot)
        frames_total += act.shape[0]
        if threshold is None:
            cols, start, vals = write.format_activations(act, classes, framehop_s, 2, time_start=chunk[0],
                                                         classes_keep=classes_keep, digits_results=digits_results)
        else:
            cols, start, vals = write.format_detections(act, threshold, classes, framehop_s, 2, chunk[0])
        _append_rows(partial, cols, start, vals)

    try:
        for i, chunk in enumerate(chunklist):
            sample_from, read_size = stream.chunk_sample_range(chunk, track.samplerate)
            track.seek(sample_from)
            pcm = np.ascontiguousarray(track.read(read_size))
            n_read = pcm.shape[0]
            short = n_read < read_size                      # bad read: the reference stops the file here
            if n_read > 0:
                if len(pending) >= n_slots:
                    drain_one()
                slot = i % n_slots
                n16 = int(engine._lib.bd_resample_out_len(n_read, track.samplerate))
                _, _, P = capi.frames_for(n16, hop_frames)
                act = np.empty((P, engine.n_classes), dtype=np.float32)
                engine.submit_pcm_ptr(slot, pcm.ctypes.data, track.fmt, track.channels, n_read, track.samplerate,
                                      hop_frames, act.ctypes.data)
                pending.append((slot, chunk, act, pcm))
            if short:
                break
        while pending:
            drain_one()
    finally:
        track.close()
    if os.path.exists(partial):
        _finalise(partial, complete)
    return {"ident": ident, "chunks": len(chunklist), "frames": frames_total, "skipped": False}


def analyze_files(paths: list[str], dir_out: str, rank: int = 0, world_size: int = 1, device: int | None = None,
                  **kw) -> list[dict]:
    """Shard whole files over ranks (buzzdetect_b200.shard) and run this rank's share on its GPU."""
    from . import shard
    from .inference.models import load_model
    durations = []
    for p in paths:
        t = open_track(p)
        durations.append(t.duration)
        t.close()
    chunklength = stream.setup_chunklength(kw.get("chunklength", 1198.08))
    plan = shard.plan([stream.file_chunklist(d, chunklength) for d in durations], world_size)
    mine = sorted({w.file_index for w in plan[rank]}) if len(paths) >= world_size else list(range(len(paths)))[rank::world_size]
    if device is not None:
        os.environ["BUZZ_B200_DEVICE"] = str(device)
    model = load_model(cfg.DEFAULT_MODEL, framehop_prop=kw.get("framehop_prop", 1.0), initialize=True)
    out = []
    for i in mine:
        out.append(analyze_wav(paths[i], dir_out, model.model, model.config["classes"], **kw))
    return out
