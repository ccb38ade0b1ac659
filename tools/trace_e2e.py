#!/usr/bin/env python
"""Timeline of the plugin end-to-end leg (bd_trace): where the wall time of `steps` recordings goes.
usage: python tools/trace_e2e.py [pcm16|f32] [steps] [out.json]"""
import json, os, sys, time, threading, queue
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("BUZZ_B200_ALLOW_SYNTHETIC", "1")
os.environ.setdefault("BUZZ_B200_SLOTS", "48")
import numpy as np
import __graft_entry__ as g
g.build()
from buzzdetect_b200 import capi
from buzzdetect_b200.inference.models import load_model
from oracle import yamnet_oracle as O

feed = sys.argv[1] if len(sys.argv) > 1 else "pcm16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
out = sys.argv[3] if len(sys.argv) > 3 else None
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


def run(steps):
    q = queue.Queue(maxsize=48)

    def writer():
        while True:
            it = q.get()
            if it is None:
                return
            it.numpy()

    th = threading.Thread(target=writer)
    th.start()
    t0 = time.perf_counter()
    for _ in range(steps):
        for (o, m) in bounds:
            q.put(model.predict_pcm(pcm16[o:o + m], SR) if feed == "pcm16" else model.predict(hv[o:o + m]))
    t_loop = time.perf_counter() - t0
    q.put(None)
    th.join()
    eng.synchronize()
    return time.perf_counter() - t0, t_loop


run(3)
run(3)
eng.trace(True)
dt, t_loop = run(steps)
recs = eng.trace(False)
print(f"{feed} steps={steps}: {steps / dt:.1f} audio-h/s, wall {dt * 1e3:.2f} ms, inferer loop {t_loop * 1e3:.2f} ms | {eng.debug_stats()}")
begin = {a: (b, h, d) for k, a, b, h, d in recs if k == 2}
end = {a: d for k, a, b, h, d in recs if k == 3}
launched = {a: h for k, a, b, h, d in recs if k == 6}
h2d = sorted(d for k, a, b, h, d in recs if k == 1)
sub = sorted(h for k, a, b, h, d in recs if k == 0)
busy = 0.0
prev_end = None
print(" pass chunks  host_launch_begin..end   dev_begin   dev_end   dur   gap_before")
for p in sorted(begin):
    b, h, d0 = begin[p]
    d1 = end.get(p, float("nan"))
    gap = d0 - prev_end if prev_end is not None else d0
    busy += d1 - d0
    print(f" {p:4d} {b:6d}  {h:9.3f} .. {launched.get(p, float('nan')):9.3f}  {d0:9.3f} {d1:9.3f} {d1 - d0:6.3f} {gap:8.3f}")
    prev_end = d1
print(f"device busy {busy:.2f} ms of {dt * 1e3:.2f} ms wall; last input copy done at {h2d[-1]:.2f} ms, first at {h2d[0]:.2f}; "
      f"submits from {sub[0]:.2f} to {sub[-1]:.2f} ms")
if out:
    json.dump({"feed": feed, "steps": steps, "wall_ms": dt * 1e3, "records": recs}, open(out, "w"))
