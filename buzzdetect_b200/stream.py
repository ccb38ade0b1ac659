"""Host-side chunk arithmetic of the streamer -- pure python/numpy, no audio I/O.

Mirrors, with the same float expressions so indices agree bit for bit:
  * Analyzer._setup_chunklength      src/analyze.py:102-111   (round to a whole number of 0.96 s frames)
  * gaps_to_chunklist                src/stream/results_coverage.py:59-70 (np.arange + round to 2 decimals)
  * melt_coverage / get_gaps / smooth_gaps  src/stream/results_coverage.py:4-56 (resume: what is still to do)
  * WorkerStreamer.queue_chunk       src/stream/worker.py:110-112 (sample_from = int(chunk[0]*sr) on python floats)

gaps_to_chunklist, get_gaps and smooth_gaps keep the reference's float expressions VERBATIM on purpose: chunk boundaries
and resume gaps must agree bit for bit with a run of the reference on the same partial file, and any re-association of
the arithmetic would move a boundary by an ulp.  They are checked against the imported reference module
(tests/test_reference_interop.py::test_resume_math_matches_reference_results_coverage); melt_coverage is a pandas-free
rewrite checked the same way.
"""
from __future__ import annotations

import numpy as np


def setup_chunklength(chunklength: float, framelength_s: float = 0.96, digits_time: int = 2) -> float:
    c = round(chunklength / framelength_s) * framelength_s
    c = round(c, digits_time)
    if c < framelength_s:
        c = framelength_s
    return c


def gaps_to_chunklist(gaps_in, chunklength, decimals=2):
    chunklist = []
    for gap in gaps_in:
        chunkpoints = np.arange(gap[0], gap[1], chunklength).tolist()
        chunkpoints.append(gap[1])
        chunkpoints = np.round(chunkpoints, decimals)
        chunklist.extend(list(zip(chunkpoints[:-1], chunkpoints[1:])))
    return chunklist


def melt_coverage(starts, framelength):
    """rows (start) of a partial result file -> merged covered intervals [(start, end)]."""
    s = np.sort(np.asarray(starts, dtype=np.float64))
    if s.size == 0:
        return []
    e = s + framelength
    prev_end = np.concatenate([[np.nan], e[:-1]])
    group = np.cumsum(s > prev_end)                  # NaN comparison is False, like pandas' shift()
    out = []
    for g in np.unique(group):
        m = group == g
        out.append((float(s[m].min()), float(e[m].max())))
    return out


def get_gaps(range_in, coverage_in):
    coverage_in = sorted(coverage_in)
    gaps = []
    if coverage_in[0][0] > range_in[0]:
        gaps.append((0, coverage_in[0][0]))
    for i in range(0, len(coverage_in) - 1):
        cur, nxt = coverage_in[i], coverage_in[i + 1]
        if nxt[0] > cur[1]:
            gaps.append((cur[1], nxt[0]))
    if coverage_in[-1][1] < range_in[1]:
        gaps.append((coverage_in[-1][1], range_in[1]))
    return gaps


def smooth_gaps(gaps, range_in, framelength, gap_tolerance):
    gaps = [g for g in gaps if g[0] < (range_in[1] - framelength)]
    if gap_tolerance is not None:
        gaps = [g for g in gaps if (g[1] - g[0]) > gap_tolerance]
    return [(g[0] - framelength / 2, g[0] + framelength / 2) if (g[1] - g[0]) < framelength else g for g in gaps]


def chunk_sample_range(chunk, samplerate: int):
    """(sample_from, read_size) exactly as queue_chunk computes them (float multiply, int() truncation)."""
    sample_from = int(chunk[0] * samplerate)
    sample_to = int(chunk[1] * samplerate)
    return sample_from, sample_to - sample_from


def file_chunklist(duration_s: float, chunklength: float, covered_starts=None, framelength_s: float = 0.96):
    """Whole file, or only the gaps left by a partial result file (resume path, src/stream/worker.py:75-106)."""
    if covered_starts is None or len(covered_starts) == 0:
        gaps = [(0, duration_s)]
    else:
        cov = melt_coverage(covered_starts, framelength_s)
        gaps = get_gaps((0, duration_s), cov)
        gaps = smooth_gaps(gaps, (0, duration_s), framelength_s, framelength_s / 4)
    return gaps_to_chunklist(gaps, chunklength)
