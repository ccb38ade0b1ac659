#!/usr/bin/env python
"""Decode the reference's own sample recording (audio_in/testbuzz.mp3: MPEG-1 Layer III, 32 kHz mono, 6.49 s) with
buzzdetect_b200.audio (FFmpeg via ctypes = the reference's PyAV fallback decoder) and store it as a 16-bit PCM WAV
fixture, so BASELINE config 1 can run where /root/reference does not exist (the GPU box).

    python tools/make_testbuzz_fixture.py [/root/reference/audio_in/testbuzz.mp3]

The fixture holds the decoded float samples rounded to int16 (the decoder's floats differ from it by <= 2^-16); the
sha256 of the exact float32 decode and of the WAV payload go to tests/golden/MANIFEST.json.
"""
import hashlib
import json
import os
import sys
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from buzzdetect_b200 import audio  # noqa: E402


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/audio_in/testbuzz.mp3"
    x, sr = audio.decode_file(src)
    assert x.ndim == 1 and sr == 32000, (x.shape, sr)
    q = np.clip(np.rint(x.astype(np.float64) * 32768.0), -32768, 32767).astype("<i2")
    out = os.path.join(ROOT, "tests", "golden", "testbuzz_32k_s16.wav")
    with wave.open(out, "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(q.tobytes())
    mpath = os.path.join(ROOT, "tests", "golden", "MANIFEST.json")
    man = json.load(open(mpath)) if os.path.exists(mpath) else {}
    man["testbuzz_32k_s16.wav"] = {
        "source": "audio_in/testbuzz.mp3 of the reference checkout", "decoder": "FFmpeg (libavcodec mp3float) via ctypes",
        "samplerate": sr, "frames": int(x.size), "float32_sha256": hashlib.sha256(x.tobytes()).hexdigest(),
        "s16_sha256": hashlib.sha256(q.tobytes()).hexdigest(),
        "max_abs_quantisation_error": float(np.abs(q.astype(np.float64) / 32768.0 - x).max()),
        "generator": "tools/make_testbuzz_fixture.py"}
    json.dump(man, open(mpath, "w"), indent=1, sort_keys=True)
    print(out, x.size, sr, man["testbuzz_32k_s16.wav"]["max_abs_quantisation_error"])


if __name__ == "__main__":
    main()
