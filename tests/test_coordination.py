"""Coordinator / manifest / log worker around per-GPU queues and the pinned-ring file
pipeline's host logic.  CPU tests use a fake model; the GPU test drives two engines on one device."""
import json
import os
import threading
import time
import wave

import numpy as np
import pytest

from buzzdetect_b200 import coordination as co
from buzzdetect_b200 import pipeline, stream

CLASSES = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "buzzdetect_b200", "assets",
                                      "config_model.json")))["classes"]


def _write_wav(path, x, sr):
    with wave.open(path, "wb") as w:
        w.setnchannels(1 if x.ndim == 1 else x.shape[1])
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(x).tobytes())


class _Lazy:
    def __init__(self, a, delay=0.0):
        self._a, self._delay = a, delay

    def numpy(self):
        if self._delay:
            time.sleep(self._delay)
        return self._a


class FakeModel:
    """Deterministic stand-in: activation[p, c] = mean |pcm| of patch p (after a trivial decimation) + c/100."""

    def __init__(self, gpu, log):
        self.gpu, self.log = gpu, log

    def predict_pcm(self, pcm, sr):
        a = np.asarray(pcm, dtype=np.float32)
        if a.ndim == 2:
            a = a.mean(axis=1)
        n16 = int(np.ceil(len(a) * (16000.0 / sr)))
        P = max(1, 1 + max(0, n16 - 15600 + 15359) // 15360)
        seg = np.array_split(np.abs(a), P)
        act = np.stack([np.full(len(CLASSES), s.mean() / 32768.0, np.float32) + np.arange(len(CLASSES), dtype=np.float32) / 100
                        for s in seg])
        self.log.append((self.gpu, len(a)))
        return _Lazy(act, 0.001)


def _factory(log):
    def f(gpu):
        return FakeModel(gpu, log)
    f.describe = {"classes": CLASSES, "framelength_s": 0.96}
    return f


# ------------------------------------------------------------------------------------------------ coordinator protocol
def test_exit_protocol_poisons_every_consumer_and_first_reason_wins():
    c = co.Coordinator(n_gpus=2, inferers_per_gpu=1, streamers_total=3)
    got = []

    def streamer():
        while c.get_stream() != co.EXIT:
            pass
        got.append("s")

    def analyzer(g):
        while c.get_analyze(g) != co.EXIT:
            pass
        got.append(f"a{g}")

    def writer():
        while c.get_write() != co.EXIT:
            pass
        got.append("w")

    ts = [threading.Thread(target=streamer) for _ in range(3)]
    ta = [[threading.Thread(target=analyzer, args=(g,))] for g in range(2)]
    tw = threading.Thread(target=writer)
    for t in ts + [t for per in ta for t in per] + [tw]:
        t.start()
    threading.Timer(0.05, c.request_stop, args=("user pressed stop",)).start()
    c.wait_for_exit(ts, ta, tw)
    for t in ts + [t for per in ta for t in per] + [tw]:
        t.join(timeout=5)
        assert not t.is_alive()
    assert sorted(got) == ["a0", "a1", "s", "s", "s", "w"]
    assert c.end_reason == "interrupted"
    c.exit_analysis(co.ExitSignal("Analysis complete", "INFO", "completed"))      # a later signal must not overwrite it
    assert c.end_reason == "interrupted"


def test_fully_analyzed_bookkeeping_and_shortest_queue_routing():
    c = co.Coordinator(n_gpus=2, depth=8)
    f = co.AssignFile("a.wav", "/tmp/out")
    chunks = [co.AssignChunk(f, (i * 10.0, i * 10.0 + 10.0), last_chunk=(i == 3)) for i in range(4)]
    for ch in chunks:
        c.put_analyze(ch)
    assert sorted(q.qsize() for q in c.q_analyze) == [2, 2]                      # balanced over the GPUs
    order = [c.get_analyze(0), c.get_analyze(1), c.get_analyze(0), c.get_analyze(1)]
    flags = []
    for ch in (order[3], order[0], order[2], order[1]):                           # written out of order
        c.put_write(ch)
        flags.append(c.get_write()[1])
    assert flags == [False, False, False, True]                                   # only the last write completes the file


def test_manifest_lock_matches_reference_semantics(tmp_path):
    d = str(tmp_path)
    m = co.build_manifest("model_general_v3", 1.0, None, ["ins_buzz", "ambient_rain"])
    assert m["output_mode"] == "activations" and m["classes_out"] == ["ambient_rain", "ins_buzz"]
    assert co.check_or_write_manifest(d, m) == (True, None)
    assert co.read_manifest(d) == m
    assert co.check_or_write_manifest(d, co.build_manifest("model_general_v3", 1.0, None, ["ambient_rain", "ins_buzz"])) == (True, None)
    ok, msg = co.check_or_write_manifest(d, co.build_manifest("model_general_v3", 0.5, None, ["ins_buzz", "mech_auto"]))
    assert not ok and "framehop_prop: existing=1.0, requested=0.5" in msg
    assert "output classes differ (added mech_auto; removed ambient_rain)" in msg
    det = co.build_manifest("model_general_v3", 1.0, 0.95, ["ins_buzz"])
    assert det["output_mode"] == "detections" and det["classes_out"] is None
    ok, msg = co.check_or_write_manifest(d, det)
    assert not ok and "output_mode" in msg and "precision" in msg


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference checkout not present")
def test_manifest_functions_agree_with_the_reference_module(tmp_path, monkeypatch):
    import importlib
    import sys
    monkeypatch.syspath_prepend("/root/reference")
    before = {k for k in sys.modules if k == "src" or k.startswith("src.")}
    try:
        ref = importlib.import_module("src.pipeline.manifest")
        cases = [("m", 1.0, None, ["b", "a"]), ("m", 0.5, 0.9, ["a"]), ("m2", 1.0, None, ["a", "c"])]
        for a in cases:
            assert co.build_manifest(*a) == ref.build_manifest(*a)
            for b in cases:
                assert co.diff_manifests(co.build_manifest(*a), co.build_manifest(*b)) == \
                    ref.diff_manifests(ref.build_manifest(*a), ref.build_manifest(*b))
        d1, d2 = str(tmp_path / "x"), str(tmp_path / "y")
        for a, b in ((cases[0], cases[0]), (cases[0], cases[1])):
            for d in (d1, d2):
                if os.path.exists(os.path.join(d, co.FNAME_MANIFEST)):
                    os.remove(os.path.join(d, co.FNAME_MANIFEST))
            co.check_or_write_manifest(d1, co.build_manifest(*a))
            ref.check_or_write_manifest(d2, ref.build_manifest(*a))
            r1 = co.check_or_write_manifest(d1, co.build_manifest(*b))
            r2 = ref.check_or_write_manifest(d2, ref.build_manifest(*b))
            assert r1[0] == r2[0] and (r1[1] or "").replace(d1, "D") == (r2[1] or "").replace(d2, "D")
    finally:
        for k in list(sys.modules):
            if (k == "src" or k.startswith("src.")) and k not in before:
                del sys.modules[k]


# ------------------------------------------------------------------------------------------------ whole driver, fake model
def test_run_analysis_two_gpus_fake_model(tmp_path):
    rng = np.random.default_rng(0)
    paths = []
    for i, secs in enumerate((25.0, 61.0, 7.5)):
        x = (rng.standard_normal(int(16000 * secs)) * 3000).astype(np.int16)
        p = str(tmp_path / f"rec{i}.wav")
        _write_wav(p, x, 16000)
        paths.append(p)
    out = str(tmp_path / "out")
    log = []
    rep = co.run_analysis(paths, out, gpus=[0, 1], chunklength=9.6, model_factory=_factory(log), streamers=3)
    assert rep["end_reason"] == "completed"
    assert {g for g, _ in log} == {0, 1}                                          # both queues were served
    for i, secs in enumerate((25.0, 61.0, 7.5)):
        f = os.path.join(out, f"rec{i}_buzzdetect.csv")
        assert os.path.exists(f) and not os.path.exists(os.path.join(out, f"rec{i}_buzzpart.csv"))
        starts = [float(l.split(",")[0]) for l in open(f).read().splitlines()[1:]]
        assert starts == sorted(starts) and len(starts) == len(set(starts))
    assert json.load(open(os.path.join(out, co.FNAME_MANIFEST)))["modelname"] == "model_general_v3"
    assert "Analysis complete" in open(os.path.join(out, "buzzdetect_b200.log")).read()
    # a second run with other settings is refused by the manifest lock; the same settings skip finished files
    with pytest.raises(ValueError, match="different settings"):
        co.run_analysis(paths, out, gpus=[0], chunklength=9.6, framehop_prop=0.5, model_factory=_factory([]))
    log2 = []
    rep2 = co.run_analysis(paths, out, gpus=[0], chunklength=9.6, model_factory=_factory(log2))
    assert rep2["end_reason"] == "completed" and log2 == []


def test_run_analysis_early_stop_leaves_resumable_partials(tmp_path):
    x = (np.random.default_rng(1).standard_normal(16000 * 400) * 3000).astype(np.int16)
    p = str(tmp_path / "long.wav")
    _write_wav(p, x, 16000)
    out = str(tmp_path / "out")

    def slow_factory(gpu):
        m = FakeModel(gpu, [])
        orig = m.predict_pcm
        m.predict_pcm = lambda pcm, sr: (time.sleep(0.03), orig(pcm, sr))[1]
        return m
    slow_factory.describe = {"classes": CLASSES, "framelength_s": 0.96}
    rep = co.run_analysis([p], out, gpus=[0], chunklength=9.6, model_factory=slow_factory, stop_after=0.25)
    assert rep["end_reason"] == "interrupted"
    assert os.path.exists(os.path.join(out, "long_buzzpart.csv")) and not os.path.exists(os.path.join(out, "long_buzzdetect.csv"))
    rep2 = co.run_analysis([p], out, gpus=[0], chunklength=9.6, model_factory=_factory([]))
    assert rep2["end_reason"] == "completed"
    lines = open(os.path.join(out, "long_buzzdetect.csv")).read().splitlines()
    starts = [float(l.split(",")[0]) for l in lines[1:]]
    assert starts == sorted(starts) and len(starts) == len(set(starts))
    assert starts[0] == 0.0 and abs(starts[-1] - 399.36) < 1e-9                   # every frame of the file exactly once


# ------------------------------------------------------------------------------------------------ pipeline host logic
def test_wav_reader_extensible_header_positioned_reads_and_framehop_guard(tmp_path):
    x = (np.random.default_rng(2).standard_normal((5000, 2)) * 4000).astype("<i2")
    p = str(tmp_path / "e.wav")
    # WAVE_FORMAT_EXTENSIBLE header written by hand (fmt chunk of 40 bytes, PCM sub-format, a LIST chunk in front of data)
    import struct
    fmt = struct.pack("<HHIIHHHHIH14s", 0xFFFE, 2, 44100, 44100 * 4, 4, 16, 22, 16, 3, 1,
                      bytes.fromhex("000000001000800000aa00389b71"))
    junk = b"LIST" + struct.pack("<I", 6) + b"abcdef"
    data = x.tobytes()
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + junk + b"data" + struct.pack("<I", len(data)) + data
    open(p, "wb").write(b"RIFF" + struct.pack("<I", len(body)) + body)
    t = pipeline.WavTrack(p)
    assert (t.samplerate, t.channels, t.frames) == (44100, 2, 5000)
    buf = np.empty((300, 2), np.int16)
    assert t.read_at(buf, 1000, 300) == 300 and np.array_equal(buf, x[1000:1300])
    assert t.read_at(buf, 4900, 300) == 100 and np.array_equal(buf[:100], x[4900:])
    t.seek(10)
    assert t.readinto(buf, 5) == 5 and np.array_equal(buf[:5], x[10:15]) and t.tell() == 15
    t.close()
    for ok in (1, 0.5, 0.25, 0.125):
        assert pipeline.validate_framehop(ok) == round(96 * ok)
    for bad in (0.3, 0.1, 0.33, 0.9):
        with pytest.raises(ValueError):
            pipeline.validate_framehop(bad)


def test_shared_partial_file_appends_and_finalise_lock(tmp_path):
    """Several ranks (one process per GPU) append to ONE partial file: the header is written once, rows are whole, and
    exactly one caller finalises."""
    part, comp = str(tmp_path / "a_buzzpart.csv"), str(tmp_path / "a_buzzdetect.csv")
    cols = ["start", "activation_x"]

    def worker(k):
        for j in range(20):
            s = np.array([k * 1000 + j * 0.96])
            pipeline._append_rows(part, cols, np.round(s, 2), np.array([[float(k)]], dtype=np.float32))

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    lines = open(part).read().splitlines()
    assert lines.count("start,activation_x") == 1 and len(lines) == 121
    assert len(pipeline._read_partial_starts(part)) == 120
    res = []
    ts = [threading.Thread(target=lambda: res.append(pipeline._finalise(part, comp))) for _ in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert sum(bool(r) for r in res) == 1 and os.path.exists(comp) and not os.path.exists(part)
    starts = [float(l.split(",")[0]) for l in open(comp).read().splitlines()[1:]]
    assert starts == sorted(starts) and len(starts) == 120


def test_make_plan_splits_one_long_file_over_ranks(tmp_path):
    x = np.zeros(16000 * 100, np.int16)
    p = str(tmp_path / "one.wav")
    _write_wav(p, x, 16000)
    plan = pipeline.make_plan([p], str(tmp_path / "o"), world_size=4, chunklength=9.6)
    sizes = [len(r) for r in plan]
    assert sum(sizes) == len(stream.file_chunklist(100.0, 9.6)) == 11 and max(sizes) - min(sizes) <= 1
    assert sorted(ci for r in plan for _, ci in r) == list(range(11))
    # at least as many files as ranks: whole files
    ps = []
    for i in range(4):
        q = str(tmp_path / f"f{i}.wav")
        _write_wav(q, x[: 16000 * (20 + 10 * i)], 16000)
        ps.append(q)
    plan = pipeline.make_plan(ps, str(tmp_path / "o"), world_size=2, chunklength=9.6)
    for r in plan:
        files = {fi for fi, _ in r}
        assert all(sum(1 for rr in plan if fi in {f for f, _ in rr}) == 1 for fi in files)


# ------------------------------------------------------------------------------------------------ GPU: the real thing
@pytest.mark.gpu
def test_bad_read_and_chunk_range_sharding_on_gpu(tmp_path, engines):
    e = engines("fp16x3", early_patches=16, late_patches=48, n_slots=8)
    rng = np.random.default_rng(5)
    sr = 32000
    x = (rng.standard_normal(sr * 50) * 3000).astype(np.int16)
    wav = str(tmp_path / "rec.wav")
    _write_wav(wav, x, sr)
    # reference run: one rank, whole file
    out_a = str(tmp_path / "a")
    r = pipeline.analyze_wav(wav, out_a, e, CLASSES, chunklength=9.6)
    full = open(os.path.join(out_a, "rec_buzzdetect.csv")).read()
    assert r["chunks"] == 6 and not r["bad_read"]
    # the same file as chunk ranges of three "ranks" (sequentially here; they share the partial file)
    out_b = str(tmp_path / "b")
    plan = pipeline.make_plan([wav], out_b, world_size=3, chunklength=9.6)
    assert [len(p) for p in plan] == [2, 2, 2]
    for rank in (2, 0, 1):
        rr = pipeline.analyze_wav(wav, out_b, e, CLASSES, chunklength=9.6, only_chunks=[plan.chunks[w] for w in plan[rank]])
        assert rr["chunks"] == 2
        assert os.path.exists(os.path.join(out_b, "rec_buzzdetect.csv")) == (rank == 1)     # the last rank finalises
    assert open(os.path.join(out_b, "rec_buzzdetect.csv")).read() == full
    # bad read: the header promises 50 s, the file holds 33.3 s (recorder died): the chunk is truncated, the file ends there
    cut = str(tmp_path / "cut.wav")
    raw = open(wav, "rb").read()
    open(cut, "wb").write(raw[: 44 + int(33.3 * sr) * 2])
    msgs = []
    t = pipeline.WavTrack(cut)
    assert t.frames == int(33.3 * sr)                    # the reader trusts the file size, as libsndfile does for RIFF
    t.close()
    # force the header's duration: emulate a driver that reports the header's frame count
    class Lying(pipeline.WavTrack):
        def _parse(self):
            super()._parse()
            self.frames_real, self.frames = self.frames, 50 * sr

        def read_at(self, out, frame, n):
            n = max(0, min(n, self.frames_real - frame))
            return super().read_at(out, frame, n) if n else 0
    orig = pipeline.open_track
    pipeline.open_track = lambda p: Lying(p)
    try:
        rb = pipeline.analyze_wav(cut, str(tmp_path / "c"), e, CLASSES, chunklength=9.6, log=lambda m, lv: msgs.append((lv, m)))
    finally:
        pipeline.open_track = orig
    assert rb["bad_read"] and rb["chunks"] == 4
    assert any(lv == "WARNING" and "Unreadable audio at 33.3s out of 50.0s" in m for lv, m in msgs)
    rows = open(os.path.join(str(tmp_path / "c"), "cut_buzzdetect.csv")).read().splitlines()
    want_rows = [l for l in full.splitlines()[1:] if float(l.split(",")[0]) < 28.8]
    assert rows[1:1 + len(want_rows)] == want_rows        # whole chunks before the bad read are identical
    assert float(rows[-1].split(",")[0]) < 33.3


@pytest.mark.gpu
def test_run_analysis_real_models_two_inferers_one_device(tmp_path, yamnet_variables, mel, head):
    """The coordinator driving the real plugin: two inferer threads (two engines) on device 0, pinned streamer ring."""
    from buzzdetect_b200 import capi
    rng = np.random.default_rng(6)
    paths = []
    for i in range(3):
        x = (rng.standard_normal(16000 * (30 + 11 * i)) * 2500).astype(np.int16)
        p = str(tmp_path / f"r{i}.wav")
        _write_wav(p, x, 16000)
        paths.append(p)
    out = str(tmp_path / "o")
    rep = co.run_analysis(paths, out, gpus=[0, 0], chunklength=9.6)
    assert rep["end_reason"] == "completed" and rep["frames"] > 0
    e = capi.Engine(device=0, allow_synthetic=True)
    try:
        ref_out = str(tmp_path / "ref")
        for p in paths:
            pipeline.analyze_wav(p, ref_out, e, CLASSES, chunklength=9.6)
    finally:
        e.close()
    for i in range(3):
        assert open(os.path.join(out, f"r{i}_buzzdetect.csv")).read() == open(os.path.join(ref_out, f"r{i}_buzzdetect.csv")).read()
