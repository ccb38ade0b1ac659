"""Writer-side semantics the activations must satisfy -- numpy only, no pandas.

Mirrors src/write/formatting.py:5-49 (add_time, format_detections, format_activations) and
src/write/thresholds.py:29-41 (calculate_threshold).  These define "identical detections": rounding to
digits_results, strict '>' against the threshold on the RAW float32 activation, start = round(i*framehop_s + t0, 2).
"""
from __future__ import annotations

import csv
import os

import numpy as np

from . import config as cfg


def frame_starts(n_frames: int, framehop_s: float, time_start: float = 0.0, digits_time: int = 2) -> np.ndarray:
    """formatting.py:5-17: range(n) * framehop_s (+ time_start if non-zero), rounded (float64, half-to-even)."""
    s = np.arange(n_frames, dtype=np.int64) * framehop_s       # int64 column * python float, as pandas does
    if time_start != 0:
        s = s + time_start
    return np.round(s, digits_time)


def format_activations(results, classes, framehop_s, digits_time, time_start=0, classes_keep='all', digits_results=2):
    """formatting.py:30-49 -> (column names, start[n], values[n, k])."""
    results = np.array(results).round(digits_results)
    classes_out = list(classes)
    if classes_keep != 'all':
        unknown = set(classes_keep) - set(classes)
        if unknown:
            raise ValueError(f"Bad classes in classes_keep: {', '.join(list(unknown))}")
        keep = [i for i, c in enumerate(classes) if c in classes_keep]
        results = results[:, keep]
        classes_out = [classes[i] for i in keep]
    cols = ['start'] + [cfg.PREFIX_COLUMN_ACTIVATION + c for c in classes_out]
    return cols, frame_starts(len(results), framehop_s, time_start, digits_time), results


def format_detections(results, threshold, classes, framehop_s, digits_time, time_start):
    """formatting.py:20-28 -> (column names, start[n], detections[n] as int)."""
    buzz_index = classes.index('ins_buzz')
    det = (np.asarray(results)[:, buzz_index] > threshold).astype(int)
    return ['start', cfg.PREFIX_COLUMN_DETECTION + 'ins_buzz'], frame_starts(len(det), framehop_s, time_start,
                                                                             digits_time), det


def calculate_threshold(modelname, precision_requested, tolerance=0.01):
    """thresholds.py:29-41: mean threshold of the metrics rows whose precision is within tolerance/2."""
    path = os.path.join(cfg.DIR_MODELS, modelname, cfg.SUBDIR_TESTS, cfg.FNAME_METRICS)
    try:
        with open(path, newline='') as f:
            rows = list(csv.DictReader(f))
    except FileNotFoundError:
        raise FileNotFoundError(
            f'metrics not available for model "{modelname}"; run test_model({modelname}) and proceed') from None
    thr = np.array([float(r['threshold']) for r in rows])
    prec = np.array([float(r['precision']) if r['precision'] not in ('', 'NA') else np.nan for r in rows])
    keep = np.abs(prec - precision_requested) <= tolerance / 2
    return float(np.mean(thr[keep])) if keep.any() else float('nan')
