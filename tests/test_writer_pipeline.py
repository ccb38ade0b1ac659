"""Writer semantics and the file-level driver.  The golden CSVs were produced by the reference's own
src/write/formatting.py + pandas.to_csv (tools/make_writer_golden.py); ours must match them byte for byte."""
import json
import os
import wave

import numpy as np
import pytest

from buzzdetect_b200 import pipeline, write

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CLASSES = json.load(open(os.path.join(os.path.dirname(GOLD), "..", "buzzdetect_b200", "assets", "config_model.json")))["classes"]


def _emit(tmp_path, name, header, start, vals):
    p = os.path.join(tmp_path, name)
    pipeline._append_rows(p, header, start, vals)
    return open(p).read()


def test_activation_csv_matches_reference_writer(tmp_path):
    act = np.load(os.path.join(GOLD, "ragged_12s.npz"))["activations"]
    for name, kw in (("writer_activations_t0", dict(time_start=0)), ("writer_activations_t199", dict(time_start=199.68))):
        cols, start, vals = write.format_activations(act, CLASSES, 0.48, 2, classes_keep="all", digits_results=2, **kw)
        assert _emit(str(tmp_path), name, cols, start, vals) == open(os.path.join(GOLD, name + ".csv")).read()
    cols, start, vals = write.format_activations(act, CLASSES, 0.48, 2, time_start=0, classes_keep=["ins_buzz", "human"])
    assert _emit(str(tmp_path), "k2", cols, start, vals) == open(os.path.join(GOLD, "writer_activations_keep2.csv")).read()
    with pytest.raises(ValueError):
        write.format_activations(act, CLASSES, 0.48, 2, classes_keep=["not_a_class"])


def test_detection_csv_matches_reference_writer(tmp_path):
    act = np.load(os.path.join(GOLD, "ragged_12s.npz"))["activations"]
    cols, start, det = write.format_detections(act, -1.2, CLASSES, 0.48, 2, 86201.28)
    assert _emit(str(tmp_path), "d", cols, start, det) == open(os.path.join(GOLD, "writer_detections.csv")).read()


def test_threshold_lookup_matches_readme():
    """models/model_general_v3/README.md:6 -- precision 0.95 <-> threshold about -1.2."""
    thr = write.calculate_threshold("model_general_v3", 0.95)
    assert -1.35 < thr < -1.05


def _write_wav(path, x_int16, sr):
    with wave.open(path, "wb") as w:
        w.setnchannels(1 if x_int16.ndim == 1 else x_int16.shape[1])
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(x_int16).tobytes())


def _synth_pcm(seconds, sr, ch, seed):
    rng = np.random.default_rng(seed)
    n = int(seconds * sr)
    t = np.arange(n) / sr
    sig = 0.25 * np.sin(2 * np.pi * 230 * t)[:, None] * (1 + np.sin(2 * np.pi * 0.3 * t))[:, None] + 0.05 * rng.standard_normal((n, ch))
    x = np.clip(sig * 16000, -32768, 32767).astype(np.int16)
    return x if ch > 1 else x[:, 0]


def test_wav_reader_and_chunk_indexing(tmp_path):
    x = _synth_pcm(3.0, 44100, 2, 0)
    p = os.path.join(str(tmp_path), "a.wav")
    _write_wav(p, x, 44100)
    t = pipeline.WavTrack(p)
    assert (t.samplerate, t.channels, t.frames) == (44100, 2, len(x))
    t.seek(1000)
    assert np.array_equal(t.read(500), x[1000:1500])
    t.seek(len(x) - 10)
    assert t.read(100).shape[0] == 10               # short read at the end of the file
    t.close()


@pytest.mark.gpu
def test_file_pipeline_resume_gives_identical_rows(tmp_path, engines, yamnet_variables, mel, head):
    """config 3 in miniature: 44.1 kHz stereo int16 WAV streamed in chunks; a run resumed from a partial file must
    produce exactly the rows of an uninterrupted run, and both must match the oracle chain."""
    from oracle import resample_oracle as R
    from oracle import yamnet_oracle as O
    e = engines("fp16x3", early_patches=16, late_patches=48)
    sr, secs = 44100, 40.0
    x = _synth_pcm(secs, sr, 2, 3)
    wav = os.path.join(str(tmp_path), "rec.wav")
    _write_wav(wav, x, sr)
    out_a = os.path.join(str(tmp_path), "a")
    r = pipeline.analyze_wav(wav, out_a, e, CLASSES, chunklength=9.6, n_in_flight=2)
    assert r["chunks"] == 5 and not os.path.exists(os.path.join(out_a, "rec_buzzpart.csv"))
    full = open(os.path.join(out_a, "rec_buzzdetect.csv")).read().splitlines()
    assert len(full) == 1 + r["frames"]
    # oracle rows for the first chunk (same float->int sample indexing, oracle resampler)
    n0 = int(9.6 * sr)
    y = R.resample(x[:n0], sr)
    a = O.predict(y, yamnet_variables, mel, head[0], head[1], 96)
    got0 = np.array([[float(v) for v in line.split(",")[1:]] for line in full[1:1 + a.shape[0]]], dtype=np.float32)
    assert np.abs(got0 - np.round(a, 2)).max() <= 0.011          # at most one rounding step (values within 1e-3)
    assert np.abs(got0 - a).max() <= 6e-3
    # resume: keep only the rows of chunks 0 and 3 as a partial file, rerun, compare
    out_b = os.path.join(str(tmp_path), "b")
    os.makedirs(out_b)
    keep = [full[0]] + [l for l in full[1:] if float(l.split(",")[0]) < 9.6 or 28.8 <= float(l.split(",")[0]) < 38.4]
    open(os.path.join(out_b, "rec_buzzpart.csv"), "w").write("\n".join(keep) + "\n")
    r2 = pipeline.analyze_wav(wav, out_b, e, CLASSES, chunklength=9.6)
    assert 0 < r2["chunks"] < 5
    resumed = open(os.path.join(out_b, "rec_buzzdetect.csv")).read().splitlines()
    assert resumed == full
    # finished files are skipped
    assert pipeline.analyze_wav(wav, out_b, e, CLASSES, chunklength=9.6)["skipped"]
