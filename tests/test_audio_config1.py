"""BASELINE config 1: the reference's own sample recording (audio_in/testbuzz.mp3, 32 kHz mono MP3) through the path.

tests/golden/testbuzz_32k_s16.wav is that file decoded with buzzdetect_b200.audio (FFmpeg via ctypes, the reference's
PyAV fallback decoder) by tools/make_testbuzz_fixture.py; the MP3 itself stays in the reference checkout."""
import hashlib
import json
import os
import wave

import numpy as np
import pytest

from buzzdetect_b200 import capi, pipeline

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WAV = os.path.join(GOLD, "testbuzz_32k_s16.wav")
MP3 = "/root/reference/audio_in/testbuzz.mp3"
CLASSES = json.load(open(os.path.join(os.path.dirname(GOLD), "..", "buzzdetect_b200", "assets", "config_model.json")))["classes"]


def _fixture():
    with wave.open(WAV, "rb") as w:
        assert (w.getframerate(), w.getnchannels(), w.getsampwidth()) == (32000, 1, 2)
        return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2")


def _ffmpeg_or_skip():
    from buzzdetect_b200 import audio
    try:
        audio._load()
    except (RuntimeError, OSError) as e:
        pytest.skip(f"no FFmpeg libraries on this box: {e}")
    return audio


def test_fixture_is_what_the_manifest_says():
    man = json.load(open(os.path.join(GOLD, "MANIFEST.json")))["testbuzz_32k_s16.wav"]
    q = _fixture()
    assert q.size == man["frames"] == 207569
    assert hashlib.sha256(q.tobytes()).hexdigest() == man["s16_sha256"]


def test_framing_of_config1():
    """6.4865 s at 32 kHz -> 103,785 samples at 16 kHz -> 7 frames starting at 0, 0.96 ... 5.76 s."""
    n16 = 103785
    assert int(np.ceil(207569 * (16000.0 / 32000))) == n16
    npad, nfr, P = capi.frames_for(n16, 96)
    assert P == 7 and npad == 15600 + 6 * 15360
    from buzzdetect_b200 import write
    cols, start, vals = write.format_activations(np.zeros((P, 13), np.float32), CLASSES, 0.96, 2, time_start=0.0)
    assert list(start) == [0.0, 0.96, 1.92, 2.88, 3.84, 4.8, 5.76]


def test_ffmpeg_decoder_reads_pcm_wav_exactly():
    """The ctypes FFmpeg path on a file every box has: the WAV fixture decodes to exactly its int16 samples / 32768."""
    audio = _ffmpeg_or_skip()
    x, sr = audio.decode_file(WAV)
    assert sr == 32000 and x.dtype == np.float32 and x.ndim == 1
    assert np.array_equal(x, _fixture().astype(np.float32) / 32768.0)
    t = pipeline.DecodedTrack(WAV)
    assert (t.samplerate, t.channels, t.frames) == (32000, 1, 207569)
    t.seek(100)
    assert np.array_equal(t.read(50), x[100:150])
    with pytest.raises(FileNotFoundError):
        audio.decode_file(WAV + ".missing")


def test_ffmpeg_decoder_keeps_channels(tmp_path):
    """Packed stereo PCM comes back as [n, 2] float32 in file order (queue_chunk downmixes afterwards)."""
    audio = _ffmpeg_or_skip()
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((4410, 2)) * 5000).astype("<i2")
    p = os.path.join(str(tmp_path), "st.wav")
    with wave.open(p, "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(2)
        w.setframerate(44100)
        w.writeframes(x.tobytes())
    y, sr = audio.decode_file(p)
    assert sr == 44100 and y.shape == (4410, 2)
    assert np.array_equal(y, x.astype(np.float32) / 32768.0)
    t = pipeline.DecodedTrack(p)
    assert t.channels == 2 and t.read(10).shape == (10, 2)


@pytest.mark.skipif(not os.path.exists(MP3), reason="reference checkout not present (GPU box)")
def test_mp3_decode_reproduces_the_fixture():
    audio = _ffmpeg_or_skip()
    man = json.load(open(os.path.join(GOLD, "MANIFEST.json")))["testbuzz_32k_s16.wav"]
    x, sr = audio.decode_file(MP3)
    assert sr == 32000 and x.shape == (207569,)
    assert hashlib.sha256(x.tobytes()).hexdigest() == man["float32_sha256"]
    q = _fixture().astype(np.float64) / 32768.0
    assert np.abs(q - x).max() <= 2.0 ** -16 + 1e-12


@pytest.mark.gpu
def test_config1_rows_match_oracle(tmp_path, engines, yamnet_variables, mel, head):
    """testbuzz through the file pipeline (int16 PCM in, device-side resample 32k -> 16k, frontend, CNN, head, CSV out)
    against the oracle chain on the same samples; timestamps exact, activations within 1e-3 + the resampler's tolerance
    (the resampler is pinned to the soxr-HQ spec, not bit-matched: DESIGN.md)."""
    from oracle import resample_oracle as R
    from oracle import yamnet_oracle as O
    e = engines("fp16x3", early_patches=16, late_patches=48)
    out = os.path.join(str(tmp_path), "o")
    r = pipeline.analyze_wav(WAV, out, e, CLASSES, chunklength=199.68)
    assert r["chunks"] == 1 and r["frames"] == 7
    lines = open(os.path.join(out, "testbuzz_32k_s16_buzzdetect.csv")).read().splitlines()
    assert lines[0].split(",") == ["start"] + ["activation_" + c for c in CLASSES]
    assert [l.split(",")[0] for l in lines[1:]] == ["0.0", "0.96", "1.92", "2.88", "3.84", "4.8", "5.76"]
    got = np.array([[float(v) for v in l.split(",")[1:]] for l in lines[1:]], dtype=np.float32)
    y = R.resample(_fixture(), 32000)
    want = O.predict(y, yamnet_variables, mel, head[0], head[1], 96)
    assert want.shape == (7, 13)
    assert np.abs(got - want).max() <= 6e-3
    # the same chunk through the float32 entry (what DecodedTrack hands over for an mp3) gives the same rows
    act = e.predict_pcm(_fixture().astype(np.float32) / 32768.0, 32000, 96)
    assert np.abs(np.round(act, 2) - got).max() <= 0.011
