#!/usr/bin/env python
"""Diagnostic 2: plugin e2e (threaded) at several step counts; run under different BD_COALESCE_PATCHES."""
import os, sys, time, threading, queue
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("BUZZ_B200_ALLOW_SYNTHETIC", "1")
os.environ.setdefault("BUZZ_B200_SLOTS", "48")
import numpy as np
import torch
import __graft_entry__ as g
g.build()
from buzzdetect_b200 import capi
from buzzdetect_b200.inference.models import load_model
from oracle import yamnet_oracle as O

SR = 16000
n = 3600 * SR
base = O.synth_audio(60 * SR, seed=1)
hv = capi.pinned_empty(n, np.float32)
for off in range(0, n, base.size):
    hv[off:off + base.size] = base[:min(base.size, n - off)]
pcm16 = capi.pinned_empty(n, np.int16)
pcm16[:] = np.clip(np.rint(hv * 32768.0), -32768, 32767).astype(np.int16)
chunk_n = int(round(199.68 * SR))
bounds = [(o, min(chunk_n, n - o)) for o in range(0, n, chunk_n)]
model = load_model("model_general_v3", framehop_prop=1, initialize=True)
eng = model.model


def run(feed, steps):
    q = queue.Queue(maxsize=48)

    def writer():
        while True:
            it = q.get()
            if it is None:
                return
            it.numpy()

    th = threading.Thread(target=writer)
    th.start()
    b0, c0 = eng.batch_stats
    t0 = time.perf_counter()
    for _ in range(steps):
        for (o, m) in bounds:
            q.put(model.predict_pcm(pcm16[o:o + m], SR) if feed == "pcm16" else model.predict(hv[o:o + m]))
    q.put(None)
    th.join()
    eng.synchronize()
    dt = time.perf_counter() - t0
    b1, c1 = eng.batch_stats
    return steps / dt, b1 - b0, (c1 - c0) / max(b1 - b0, 1)


for feed in ("pcm16", "f32"):
    run(feed, 3)
    for steps in (5, 5, 20):
        v, p, cp = run(feed, steps)
        print(f"target={os.environ.get('BD_COALESCE_PATCHES','default')} {feed} steps={steps}: {v:7.1f} audio-h/s, {p} passes, {cp:.1f} chunks/pass | {eng.debug_stats()}", flush=True)
