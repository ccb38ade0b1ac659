"""Run-time probe for the REAL reference runtime.

The reference's arithmetic lives in TensorFlow (tf.signal.stft, Conv2D, ...) and librosa/soxr, none of which can be
installed in the build image.  On a box that has them, true parity can be measured instead of parity against the
restated oracle: bench.py records what this probe finds, and tests/test_real_reference_probe.py runs the real frontend
ops and the real resampler against the CUDA path when they import.  Nothing here is on the product path."""
from __future__ import annotations

import importlib


def _version(name: str):
    try:
        m = importlib.import_module(name)
    except Exception:                                          # ImportError, or a broken install
        return None
    return getattr(m, "__version__", "unknown")


def reference_runtime() -> dict:
    r = {k: _version(k) for k in ("tensorflow", "librosa", "soxr")}
    r["complete"] = all(r[k] is not None for k in ("tensorflow", "librosa", "soxr"))
    return r


def tf_log_mel(samples_padded, mel_matrix):
    """embedders/yamnet/features.py:42-58 with real TensorFlow ops (None when TensorFlow is absent)."""
    try:
        import tensorflow as tf
    except Exception:
        return None
    import numpy as np
    x = tf.convert_to_tensor(np.asarray(samples_padded, dtype=np.float32))
    stft = tf.signal.stft(signals=x, frame_length=400, frame_step=160, fft_length=512)
    mag = tf.abs(stft)
    mel = tf.matmul(mag, tf.convert_to_tensor(np.asarray(mel_matrix, dtype=np.float32)))
    return tf.math.log(mel + 0.001).numpy()


def librosa_resample(samples, orig_sr: int, target_sr: int = 16000):
    """src/stream/worker.py:128 with the real librosa / soxr (None when absent)."""
    try:
        import librosa
    except Exception:
        return None
    import numpy as np
    return librosa.resample(y=np.asarray(samples, dtype=np.float32), orig_sr=orig_sr, target_sr=target_sr)
