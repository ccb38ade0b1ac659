"""If the box has the real reference runtime (TensorFlow + librosa + soxr), measure TRUE parity of the two stages
whose arithmetic lives in those libraries; otherwise record that it is absent (it is, in the build image and on the
GPU pool's image).  The SavedModel graphs themselves cannot run anywhere the YAMNet variables blob is missing."""
import numpy as np
import pytest

from buzzdetect_b200 import probe
from oracle import yamnet_oracle as O

pytestmark = pytest.mark.gpu


def test_true_parity_when_the_reference_runtime_is_present(engines, mel, parity_report):
    rt = probe.reference_runtime()
    parity_report("reference_runtime", rt)
    if not rt["complete"]:
        pytest.skip(f"reference runtime absent on this box: {rt}")
    e = engines("fp32")
    x = O.synth_audio(16000 * 10, seed=21)
    xp = O.pad_waveform(x, 96)
    nf = 1 + (len(xp) - 400) // 160
    want = probe.tf_log_mel(xp, mel)
    got = e.debug_logmel(x, nf)
    err = float(np.abs(got - want[:nf]).max())
    parity_report("true_parity_logmel_vs_tensorflow", err)
    assert err <= 1e-4
    src = O.synth_audio(44100 * 5, seed=22, sr=44100)
    want = probe.librosa_resample(src, 44100)
    got = e.resample(src, 44100)
    assert got.shape == want.shape
    delta = float(np.abs(got - want).max())
    parity_report("true_parity_resample_vs_soxr", delta)
    assert delta <= 1e-3
