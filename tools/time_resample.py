#!/usr/bin/env python
"""Device time of the downmix + resample stage: one chunk of `--seconds` at `--rate` Hz with `--channels` int16 channels
through bd_submit_pcm_host, against the same audio already at 16 kHz mono (difference = resampler + wider H2D)."""
import argparse
import sys
import time
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=599.04)
    ap.add_argument("--rate", type=int, default=44100)
    ap.add_argument("--channels", type=int, default=2)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    import __graft_entry__ as g
    g.build()
    from buzzdetect_b200 import capi
    eng = capi.Engine(device=0, n_slots=2)
    n = int(a.seconds * a.rate)
    rng = np.random.default_rng(0)
    pcm = torch.from_numpy((rng.standard_normal((n, a.channels)) * 3000).astype(np.int16)).pin_memory()
    n16 = int(eng._lib.bd_resample_out_len(n, a.rate))
    _, _, P = capi.frames_for(n16, 96)
    act = torch.empty((P, eng.n_classes), dtype=torch.float32).pin_memory()
    x16 = torch.from_numpy(rng.standard_normal(n16).astype(np.float32) * 0.05).pin_memory()

    def run_pcm():
        eng.submit_pcm_ptr(0, pcm.data_ptr(), 1, a.channels, n, a.rate, 96, act.data_ptr())
        eng.wait(0)

    def run_16k():
        eng.submit_ptr(0, x16.data_ptr(), n16, 96, act.data_ptr())
        eng.wait(0)

    out = {}
    for name, fn in (("pcm", run_pcm), ("mono16k", run_16k)):
        for _ in range(2):
            fn()
        t0 = time.perf_counter()
        for _ in range(a.reps):
            fn()
        out[name] = (time.perf_counter() - t0) / a.reps * 1e3
    hours = a.seconds / 3600.0
    print(f"{a.rate} Hz x{a.channels} int16, {a.seconds} s chunk: pcm path {out['pcm']:.3f} ms, 16 kHz mono path "
          f"{out['mono16k']:.3f} ms -> resample stage ~{(out['pcm'] - out['mono16k']) / hours:.2f} ms per audio-hour "
          f"(pcm bytes {pcm.numel() * 2 / 1e6:.0f} MB, mono bytes {x16.numel() * 4 / 1e6:.0f} MB)")
    eng.close()


if __name__ == "__main__":
    main()
