#!/usr/bin/env python
"""Diagnostic: where does the plugin end-to-end leg spend its time?  (GPU box only)"""
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


def run(feed, steps, threaded=True):
    q = queue.Queue(maxsize=48)
    tcall = []

    def writer():
        while True:
            it = q.get()
            if it is None:
                return
            it.numpy()

    th = threading.Thread(target=writer)
    if threaded:
        th.start()
    held = []
    b0, c0 = eng.batch_stats
    t0 = time.perf_counter()
    for _ in range(steps):
        for (o, m) in bounds:
            t1 = time.perf_counter()
            r = model.predict_pcm(pcm16[o:o + m], SR) if feed == "pcm16" else model.predict(hv[o:o + m])
            tcall.append(time.perf_counter() - t1)
            if threaded:
                q.put(r)
            else:
                held.append(r)
                if len(held) >= 40:
                    held.pop(0).numpy()
    t_loop = time.perf_counter() - t0
    if threaded:
        q.put(None)
        th.join()
    else:
        for r in held:
            r.numpy()
    eng.synchronize()
    dt = time.perf_counter() - t0
    b1, c1 = eng.batch_stats
    tc = np.array(tcall) * 1e3
    print(f"{feed:6s} threaded={threaded} steps={steps}: {steps / dt:7.1f} audio-h/s  wall {dt*1e3:7.1f} ms  loop {t_loop*1e3:7.1f} ms  "
          f"passes {b1-b0} ({(c1-c0)/max(b1-b0,1):.1f} chunks/pass)  predict call ms: median {np.median(tc):.3f} p90 {np.percentile(tc,90):.3f} max {tc.max():.3f}",
          flush=True)


for rep in range(2):
    for feed in ("pcm16", "f32"):
        for threaded in (True, False):
            run(feed, 3, threaded)
# raw engine timing of one coalesced pass of 16 chunks, by hand
eng.set_auto_flush(False)
for feed in ("pcm16", "f32"):
    for rep in range(3):
        t0 = time.perf_counter()
        tks = [eng.submit_pcm(pcm16[o:o + m], SR, 96) if feed == "pcm16" else eng.submit(hv[o:o + m], 96) for (o, m) in bounds[:16]]
        t1 = time.perf_counter()
        eng.flush()
        t2 = time.perf_counter()
        for t in tks:
            t.result()
        t3 = time.perf_counter()
        print(f"manual {feed}: submit 16 chunks {1e3*(t1-t0):.2f} ms, flush {1e3*(t2-t1):.2f} ms, results {1e3*(t3-t2):.2f} ms", flush=True)
eng.set_auto_flush(True)
