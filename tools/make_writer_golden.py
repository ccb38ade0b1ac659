#!/usr/bin/env python
"""tests/golden/writer_*.csv: what the REFERENCE's own writer code emits (src/write/formatting.py + pandas to_csv, as
called by src/write/worker.py:67-81) for fixed activations.  formatting.py only needs numpy + pandas, so it is
imported from /root/reference directly.  Run in the build container."""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = os.environ.get("BUZZ_REFERENCE", "/root/reference")
spec = importlib.util.spec_from_file_location("ref_formatting", os.path.join(REF, "src/write/formatting.py"))
fmt = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fmt)

classes = json.load(open(os.path.join(REF, "models/model_general_v3/config_model.json")))["classes"]
g = np.load(os.path.join(ROOT, "tests/golden/ragged_12s.npz"))
act = g["activations"]
out = os.path.join(ROOT, "tests/golden")
for name, kw in (("writer_activations_t0", dict(time_start=0)), ("writer_activations_t199", dict(time_start=199.68))):
    df = fmt.format_activations(act, classes=classes, framehop_s=0.48, digits_time=2, classes_keep='all',
                                digits_results=2, **kw)
    df.to_csv(os.path.join(out, name + ".csv"), index=False)
df = fmt.format_activations(act, classes=classes, framehop_s=0.48, digits_time=2, time_start=0,
                            classes_keep=['ins_buzz', 'human'], digits_results=2)
df.to_csv(os.path.join(out, "writer_activations_keep2.csv"), index=False)
df = fmt.format_detections(act, threshold=-1.2, classes=classes, framehop_s=0.48, digits_time=2, time_start=86201.28)
df.to_csv(os.path.join(out, "writer_detections.csv"), index=False)
print(open(os.path.join(out, "writer_activations_t199.csv")).read()[:600])
print(open(os.path.join(out, "writer_detections.csv")).read()[:200])
