"""CPU ORACLE (test infrastructure) for the downmix + resample step of the streamer.

Reference call site: src/stream/worker.py:116-117 (np.mean over channels in float32) and :128
(librosa.resample(y, orig_sr=sr, target_sr=16000), default res_type 'soxr_hq').  librosa and soxr are third-party
libraries that are NOT in /root/reference (environment.yml:9 leaves both unpinned) and cannot be installed here, so
bit parity with soxr is impossible by construction -- "parity unpinned" for this step.  What can be pinned, and is:

  * output length: int(ceil(n * 16000/sr)) (librosa.resample -> fix_length), identity when sr == 16000;
  * zero state per chunk, linear phase with the delay compensated (output sample m sits at input time m*sr/16000);
  * the published soxr "HQ" band spec: pass band to 0.9136 of the lower Nyquist, stop band from that Nyquist,
    >= 120 dB rejection -- evaluated here with a Kaiser-windowed sinc (beta from the 125 dB design rule).

This file evaluates that interpolation in float64 straight from the continuous kernel (no polyphase table), which
is the restatement the CUDA kernel (csrc/resample.cu, taps from engine.cu:get_resampler) is compared against.
"""
from __future__ import annotations

import math

import numpy as np

TARGET = 16000
ATT_DB = 125.0
PASS_FRAC = 0.9136


def out_len(n: int, sr: int) -> int:
    if n <= 0:
        return 0
    if sr == TARGET:
        return n
    return int(math.ceil(n * (float(TARGET) / sr)))


def downmix(x: np.ndarray) -> np.ndarray:
    """np.mean(samples, axis=1) on float32 [n, C] (int16 input is first scaled by 1/32768 like soundfile)."""
    a = np.asarray(x)
    if a.dtype == np.int16:
        a = a.astype(np.float32) * np.float32(1.0 / 32768.0)
    a = a.astype(np.float32, copy=False)
    if a.ndim == 1:
        return a
    s = np.zeros(a.shape[0], dtype=np.float32)
    for c in range(a.shape[1]):
        s = s + a[:, c]
    return (s / np.float32(a.shape[1])).astype(np.float32)


def design(sr: int):
    g = math.gcd(TARGET, sr)
    up, down = TARGET // g, sr // g
    f_low = 0.5 * min(TARGET, sr)
    fpass, fstop = PASS_FRAC * f_low, f_low
    fc = 0.5 * (fpass + fstop)
    beta = 0.1102 * (ATT_DB - 8.7)
    fs_v = float(up) * sr
    dw = 2.0 * math.pi * (fstop - fpass) / fs_v
    half = int(math.ceil((ATT_DB - 8.0) / (2.285 * dw) / 2.0))
    return up, down, fc, beta, fs_v, half


def kernel(t_virtual: np.ndarray, sr: int) -> np.ndarray:
    """Prototype low-pass h(t) at virtual-sample offsets t (float64), gain `up` in the pass band."""
    up, down, fc, beta, fs_v, half = design(sr)
    r = t_virtual / half
    inside = np.abs(r) <= 1.0
    arg = 2.0 * fc / fs_v * t_virtual
    sinc = np.sinc(arg)
    win = np.where(inside, np.i0(beta * np.sqrt(np.clip(1.0 - r * r, 0.0, 1.0))) / np.i0(beta), 0.0)
    return np.where(inside, 2.0 * fc / fs_v * sinc * win * up, 0.0)


def resample(x: np.ndarray, sr: int) -> np.ndarray:
    """float64 evaluation: y[m] = sum_n x[n] h(m*down - n*up)."""
    x = downmix(x)
    n = len(x)
    no = out_len(n, sr)
    if sr == TARGET:
        return x.astype(np.float32)
    up, down, fc, beta, fs_v, half = design(sr)
    y = np.zeros(no, dtype=np.float64)
    xd = x.astype(np.float64)
    span = half // up + 2
    for m in range(no):
        v = m * down
        n0 = v // up
        lo, hi = max(0, n0 - span), min(n - 1, n0 + span)
        if hi < lo:
            continue
        idx = np.arange(lo, hi + 1)
        t = (v - idx * up).astype(np.float64)
        y[m] = np.dot(xd[idx], kernel(t, sr))
    return y.astype(np.float32)


def frequency_response(sr: int, freqs_hz: np.ndarray) -> np.ndarray:
    """|H(f)| of the prototype (relative to pass-band gain 1), evaluated at the virtual rate."""
    up, down, fc, beta, fs_v, half = design(sr)
    t = np.arange(-half, half + 1, dtype=np.float64)
    h = kernel(t, sr) / up
    w = 2.0 * np.pi * np.asarray(freqs_hz, dtype=np.float64)[:, None] / fs_v
    return np.abs((h[None, :] * np.exp(-1j * w * t[None, :])).sum(axis=1))
