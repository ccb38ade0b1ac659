"""CPU tests against the reference's OWN importable modules (only where /root/reference exists: this container, not
the GPU box) and against a second, independent sample-rate converter.

  * stream.py's resume arithmetic vs /root/reference/src/stream/results_coverage.py (pandas) on randomised partial files
  * the plugin files loaded through the reference's own src/inference/models.load_model from a checkout-shaped tree
  * how much the activations depend on the resampling filter: our HQ-spec filter vs torchaudio's Kaiser-windowed sinc
    (soxr itself is not in this image)
  * the run-time probe for the real reference runtime (TensorFlow / librosa / soxr)
"""
import importlib
import os
import shutil
import sys

import numpy as np
import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not present")


@pytest.fixture
def reference_on_path(monkeypatch):
    """/root/reference importable as `src...` for one test, with every trace removed afterwards."""
    before = {k for k in sys.modules if k == "src" or k.startswith("src.")}
    monkeypatch.syspath_prepend(REF)
    importlib.invalidate_caches()
    yield
    for k in list(sys.modules):
        if (k == "src" or k.startswith("src.")) and k not in before:
            del sys.modules[k]


# ------------------------------------------------------------------------------------------------ resume math
@needs_ref
def test_resume_math_matches_reference_results_coverage(reference_on_path):
    import pandas as pd
    rc = importlib.import_module("src.stream.results_coverage")
    from buzzdetect_b200 import stream
    rng = np.random.default_rng(7)
    framelength = 0.96
    for trial in range(200):
        duration = float(rng.uniform(5, 4000))
        hop = float(rng.choice([0.96, 0.48]))
        n_frames = int(duration // hop)
        starts = np.round(np.arange(n_frames) * hop, 2)
        # knock out a few random stretches (interrupted runs), sometimes nothing, sometimes almost everything
        keep = np.ones(n_frames, dtype=bool)
        for _ in range(int(rng.integers(0, 5))):
            a = int(rng.integers(0, max(n_frames, 1)))
            b = a + int(rng.integers(1, max(2, n_frames // 3 + 1)))
            keep[a:b] = False
        if trial % 17 == 0:
            keep[:] = False
            keep[int(rng.integers(0, max(n_frames, 1))) % max(n_frames, 1)] = True
        covered = starts[keep]
        rng.shuffle(covered)                                 # partial files are appended out of order by several inferers
        if covered.size == 0:
            continue
        df = pd.DataFrame({"start": covered, "activation_ins_buzz": 0.0})
        chunklength = stream.setup_chunklength(float(rng.choice([199.68, 1198.08, 30.0, 0.5])), framelength)
        # the reference's sequence: src/stream/worker.py:85-106
        cov_ref = rc.melt_coverage(df, framelength)
        gaps_ref = rc.get_gaps(range_in=(0, duration), coverage_in=cov_ref)
        gaps_ref = rc.smooth_gaps(gaps_ref, range_in=(0, duration), framelength=framelength, gap_tolerance=framelength / 4)
        chunks_ref = rc.gaps_to_chunklist(gaps_ref, chunklength)
        cov = stream.melt_coverage(covered, framelength)
        assert [(float(a), float(b)) for a, b in cov] == [(float(a), float(b)) for a, b in cov_ref], trial
        chunks = stream.file_chunklist(duration, chunklength, covered, framelength)
        assert [(float(a), float(b)) for a, b in chunks] == [(float(a), float(b)) for a, b in chunks_ref], trial
        # and the sample ranges the streamer would read (python-float multiply + int() truncation)
        for c in chunks[:5]:
            sf, rs = stream.chunk_sample_range(c, 44100)
            assert sf == int(c[0] * 44100) and rs == int(c[1] * 44100) - int(c[0] * 44100)
    # whole-file chunking and the chunk-length rounding of Analyzer._setup_chunklength (src/analyze.py:102-111)
    for cl in (200, 199.68, 1200, 0.3, 37.123):
        c = stream.setup_chunklength(cl, framelength)
        assert c == max(round(round(cl / framelength) * framelength, 2), framelength)
        ref = rc.gaps_to_chunklist([(0, 3601.5)], c)
        got = stream.file_chunklist(3601.5, c)
        assert [(float(a), float(b)) for a, b in got] == [(float(a), float(b)) for a, b in ref]


# ------------------------------------------------------------------------------------------------ reference loader
@needs_ref
def test_plugins_load_through_the_reference_loader(reference_on_path, tmp_path, monkeypatch):
    """A checkout-shaped tree (embedders/<name>/embedder.py, models/<name>/{model.py, config_model.json}) holding OUR
    plugin files, driven by the reference's own src.inference.models.load_model (DIR_MODELS / DIR_EMBEDDERS are
    relative to the CWD, src/config.py:23-26).  initialize=False: no GPU needed."""
    plug = os.path.join(ROOT, "buzzdetect_b200", "plugins")
    for name in ("yamnet", "yamnet_k2"):
        os.makedirs(tmp_path / "embedders" / name)
        shutil.copy(os.path.join(plug, "embedders", name, "embedder.py"), tmp_path / "embedders" / name / "embedder.py")
    os.makedirs(tmp_path / "models" / "model_general_v3")
    shutil.copy(os.path.join(plug, "models", "model_general_v3", "model.py"), tmp_path / "models" / "model_general_v3" / "model.py")
    shutil.copy(os.path.join(REF, "models", "model_general_v3", "config_model.json"),
                tmp_path / "models" / "model_general_v3" / "config_model.json")
    monkeypatch.chdir(tmp_path)
    ref_models = importlib.import_module("src.inference.models")
    ref_embedding = importlib.import_module("src.inference.embedding")
    assert ref_models.__file__.startswith(REF)
    model = ref_models.load_model("model_general_v3", framehop_prop=1, initialize=False)
    assert isinstance(model, ref_models.BaseModel) and type(model).__name__ == "ModelGeneralV3"
    assert isinstance(model.embedder, ref_embedding.BaseEmbedder) and type(model.embedder).__name__ == "YamnetK2"
    # the attributes Analyzer / WorkerStreamer / WorkerWriter read BEFORE initialize()
    emb = model.embedder
    assert (emb.samplerate, emb.framelength_s, emb.n_embeddings, emb.digits_time, emb.dtype_in) == (16000, 0.96, 1024, 2, "float32")
    assert emb.framehop_prop == 1 and emb.framehop_s == 0.96 and emb.model is None and model.model is None
    assert model.modelname == "model_general_v3" and model.digits_results == 2
    assert model.config["classes"][8] == "ins_buzz" and len(model.config["classes"]) == 13
    # same values as the reference's own plugin classes declare
    ref_plugin = {}
    for name in ("yamnet", "yamnet_k2"):
        src_ = open(os.path.join(REF, "embedders", name, "embedder.py")).read()
        for attr in ("samplerate", "framelength_s", "n_embeddings", "digits_time"):
            assert f"{attr} = {getattr(emb, attr)}" in src_, (name, attr)
    e3 = ref_embedding.load_embedder("yamnet", framehop_prop=0.5, initialize=False)
    assert type(e3).__name__ == "EmbedderYamnet" and e3.framehop_s == 0.48
    # the override hook: same head over the Keras-3 embedder (reference model.py:8 hard-codes yamnet_k2)
    monkeypatch.setenv("BUZZ_B200_EMBEDDER", "yamnet")
    m3 = ref_models.load_model("model_general_v3", framehop_prop=1, initialize=False)
    assert type(m3.embedder).__name__ == "EmbedderYamnet"
    with pytest.raises(ValueError):
        ref_models.load_model("no_such_model", framehop_prop=1, initialize=False)


def test_production_path_refuses_synthetic_weights(monkeypatch, tmp_path):
    """Without the real blob and without the explicit opt-in, weight resolution fails loudly."""
    from buzzdetect_b200 import weights as W
    monkeypatch.delenv("BUZZ_B200_ALLOW_SYNTHETIC", raising=False)
    monkeypatch.delenv("BUZZ_YAMNET_WEIGHTS", raising=False)
    monkeypatch.delenv("BUZZDETECT_ROOT", raising=False)
    monkeypatch.chdir(tmp_path)
    with pytest.raises(FileNotFoundError, match="ALLOW_SYNTHETIC"):
        W.resolve_yamnet()
    v, prov = W.resolve_yamnet(allow_synthetic=True)
    assert prov.startswith("synthetic:") and len(v) == 108                # same tensor names as the checkpoint index
    monkeypatch.setenv("BUZZ_B200_ALLOW_SYNTHETIC", "1")
    assert W.resolve_yamnet()[1].startswith("synthetic:")


# ------------------------------------------------------------------------------------------------ filter dependence
@pytest.mark.parametrize("sr", [32000, 44100])
def test_activation_delta_between_two_resampling_filters(sr, yamnet_variables, mel, head):
    """soxr cannot be run here, so the resampler is pinned to the HQ spec, not to soxr's taps.  This bounds what that
    costs downstream: the same audio resampled by OUR filter (float64 oracle evaluation) and by an independent
    converter (torchaudio sinc_interp_kaiser, different pass band / roll-off) -- activation differences through the
    oracle network on the seeded synthetic weights.  Recorded in DESIGN.md section 4."""
    import torch
    import torchaudio.functional as AF
    from oracle import resample_oracle as R
    from oracle import yamnet_oracle as O
    secs = 4.0
    x = O.synth_audio(int(sr * secs), seed=31, sr=sr)
    ours = R.resample(x, sr)
    theirs = AF.resample(torch.from_numpy(x)[None], sr, 16000, lowpass_filter_width=64, rolloff=0.9475937167399596,
                         resampling_method="sinc_interp_kaiser", beta=14.769656459379492)[0].numpy()
    assert len(ours) == R.out_len(len(x), sr)
    k = min(len(ours), len(theirs))
    assert abs(len(ours) - len(theirs)) <= 1
    a = O.predict(ours[:k], yamnet_variables, mel, head[0], head[1], 96)
    b = O.predict(theirs[:k].astype(np.float32), yamnet_variables, mel, head[0], head[1], 96)
    wave_delta = float(np.abs(ours[:k] - theirs[:k]).max())
    act_delta = float(np.abs(a - b).max())
    flips = int(((a[:, 8] > -1.2) != (b[:, 8] > -1.2)).sum())
    print(f"resample {sr}->16000: waveform max delta {wave_delta:.3e}, activation max delta {act_delta:.3e}, flips {flips}")
    assert wave_delta < 0.05
    assert act_delta < 0.05


# ------------------------------------------------------------------------------------------------ runtime probe
def test_reference_runtime_probe_reports_what_is_installed():
    from buzzdetect_b200 import probe
    r = probe.reference_runtime()
    assert set(r) >= {"tensorflow", "librosa", "soxr", "complete"}
    for k in ("tensorflow", "librosa", "soxr"):
        try:
            importlib.import_module(k)
            assert r[k] is not None
        except ImportError:
            assert r[k] is None
    assert r["complete"] == all(r[k] is not None for k in ("tensorflow", "librosa", "soxr"))
