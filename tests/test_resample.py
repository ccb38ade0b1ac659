"""Downmix + resample (src/stream/worker.py:116-128).  soxr cannot be bit-matched (third-party, absent): the oracle
is pinned to the published HQ spec and to librosa's length rule; the CUDA kernel is compared with the oracle."""
import numpy as np
import pytest

from oracle import resample_oracle as R


@pytest.mark.parametrize("sr", [44100, 48000, 32000, 22050, 8000, 96000])
def test_filter_meets_hq_spec(sr):
    f_low = 0.5 * min(16000, sr)
    fp = np.linspace(0, R.PASS_FRAC * f_low, 200)
    fs = np.linspace(f_low, min(4 * f_low, 0.5 * R.design(sr)[4]), 400)
    hp = R.frequency_response(sr, fp)
    hs = R.frequency_response(sr, fs)
    assert np.abs(20 * np.log10(hp)).max() < 0.01            # pass-band ripple < 0.01 dB
    assert (20 * np.log10(hs + 1e-30)).max() < -120.0        # >= 120 dB rejection from the lower Nyquist on


@pytest.mark.parametrize("n,sr,expect", [(8805888, 44100, 3194880), (209664, 32000, 104832), (1, 44100, 1),
                                         (441, 44100, 160), (442, 44100, 161), (1000, 16000, 1000), (0, 44100, 0),
                                         (16000, 8000, 32000)])
def test_output_length_rule(n, sr, expect, built_lib):
    assert R.out_len(n, sr) == expect
    assert built_lib.bd_resample_out_len(n, sr) == expect


def test_zero_phase_and_unity_gain():
    sr = 44100
    t = np.arange(sr // 2) / sr
    x = np.sin(2 * np.pi * 1000.0 * t).astype(np.float32)
    y = R.resample(x, sr)
    tt = np.arange(len(y)) / 16000.0
    ref = np.sin(2 * np.pi * 1000.0 * tt)
    mid = slice(400, len(y) - 400)                            # away from the zero-state edges
    assert np.abs(y[mid] - ref[mid]).max() < 1e-5


def test_downmix_matches_numpy_mean():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((1000, 2)).astype(np.float32)
    assert np.array_equal(R.downmix(x), np.mean(x, axis=1))
    xi = (rng.standard_normal((1000, 2)) * 8000).astype(np.int16)
    assert np.array_equal(R.downmix(xi), np.mean(xi.astype(np.float32) / np.float32(32768.0), axis=1))


NO_TC = 1 << 21        # BD_FUSE_NO_TC_RESAMPLE: tap-by-tap CUDA-core evaluation instead of the tcgen05 GEMM


@pytest.mark.gpu
@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("sr,ch,dtype", [(44100, 2, np.int16), (44100, 1, np.float32), (32000, 1, np.float32),
                                         (48000, 2, np.float32), (22050, 1, np.int16), (8000, 1, np.float32),
                                         (16000, 2, np.int16), (16000, 1, np.float32), (96000, 3, np.int16),
                                         (11025, 1, np.float32), (24000, 2, np.int16)])
def test_cuda_resampler_matches_oracle(engines, sr, ch, dtype, tc):
    e = engines("fp32") if tc else engines("fp32", fuse_mask=NO_TC)
    rng = np.random.default_rng(sr + ch)
    n = int(sr * (2.35 if tc else 0.35)) + 17                  # the GEMM path needs several whole blocks to be exercised
    t = np.arange(n) / sr
    sig = 0.4 * np.sin(2 * np.pi * 440.0 * t)[:, None] + 0.2 * rng.standard_normal((n, ch))
    if dtype == np.int16:
        x = np.clip(sig * 20000, -32768, 32767).astype(np.int16)
    else:
        x = sig.astype(np.float32)
    if ch == 1:
        x = x[:, 0]
    got = e.resample(x, sr)
    want = R.resample(x, sr)
    assert got.shape == want.shape == (R.out_len(n, sr),)
    if sr == 16000:
        assert np.array_equal(got, want)                      # identity / pure downmix: bit exact
    else:
        err = float(np.abs(got - want).max())
        assert err < 5e-6, err


@pytest.mark.gpu
@pytest.mark.parametrize("sr,ch", [(12000, 1), (37800, 2), (88200, 2), (192000, 1), (6000, 1), (47999, 1)])
def test_unusual_rates_both_evaluations_agree(engines, sr, ch):
    """Rates with other up/down factors (block sizes 32..160 x n-tiles, up-sampling, and one ratio whose interpolation
    factor is too large for either table: both paths must refuse it the same way)."""
    rng = np.random.default_rng(sr)
    n = int(sr * 1.7) + 5
    x = (rng.standard_normal((n, ch)) * 0.2).astype(np.float32)
    x = x[:, 0] if ch == 1 else x
    if sr == 47999:
        with pytest.raises(RuntimeError, match="unsupported sample-rate ratio"):
            engines("fp32").resample(x, sr)
        return
    a = engines("fp32").resample(x, sr)
    b = engines("fp32", fuse_mask=NO_TC).resample(x, sr)
    assert a.shape == b.shape == (R.out_len(n, sr),)
    assert np.isfinite(a).all() and float(np.abs(a - b).max()) < 2e-6


@pytest.mark.gpu
def test_tensor_core_resampler_equals_tap_by_tap_kernel(engines):
    """Same filter, two evaluations: the GEMM (fp16 hi/lo split, fp32 accumulate) against the CUDA-core loop on 20 s of
    44.1 kHz stereo int16 -- including the first block (zero state before the chunk) and the tail after the last whole
    block, which the GEMM path hands to the tap-by-tap kernel."""
    rng = np.random.default_rng(9)
    n = 44100 * 20 + 123
    x = np.clip(rng.standard_normal((n, 2)) * 6000, -32768, 32767).astype(np.int16)
    a = engines("fp32").resample(x, 44100)
    b = engines("fp32", fuse_mask=NO_TC).resample(x, 44100)
    assert a.shape == b.shape
    assert float(np.abs(a - b).max()) < 2e-6
    assert np.array_equal(a[-(a.size % 160):], b[-(b.size % 160):])        # the tail is the same kernel


@pytest.mark.gpu
def test_resampled_audio_through_the_path(engines, yamnet_variables, mel, head):
    """44.1 kHz stereo int16 -> resample on the GPU -> predict: activations within tolerance of the oracle chain."""
    from oracle import yamnet_oracle as O
    e = engines("fp16x3", early_patches=16, late_patches=48)
    rng = np.random.default_rng(5)
    n = 44100 * 6
    x = np.clip(rng.standard_normal((n, 2)) * 1500 + 4000 * np.sin(2 * np.pi * 250 * np.arange(n) / 44100)[:, None],
                -32768, 32767).astype(np.int16)
    y_gpu = e.resample(x, 44100)
    y_ref = R.resample(x, 44100)
    a_gpu = e.predict(y_gpu, 96)
    a_ref = O.predict(y_ref, yamnet_variables, mel, head[0], head[1], 96)
    assert a_gpu.shape == a_ref.shape
    assert np.abs(a_gpu - a_ref).max() <= 1e-3
