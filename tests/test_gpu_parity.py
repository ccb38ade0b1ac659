"""GPU parity tests proper: every call goes through the C ABI (ctypes) and is compared with the CPU oracle.

Tolerances (BASELINE.json north_star): frame indexing exact; activations max-abs error <= 1e-3 with identical
thresholded detections for the default precision (fp16x3) and the fp32 SIMT mode; the single-pass fp16 mode is a
documented faster / looser option (<= 5e-3 of the activation range)."""
import json
import os

import numpy as np
import pytest

from oracle import yamnet_oracle as O

pytestmark = pytest.mark.gpu

REPORT = {}


def _report(key, val):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_report.json")
        if not REPORT and os.path.exists(path):
            with open(path) as f:
                REPORT.update(json.load(f))          # several pytest invocations append to one report
        REPORT[key] = val
        with open(path, "w") as f:
            json.dump(REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


# ------------------------------------------------------------------------------------------------ frontend
@pytest.mark.parametrize("n", [15600, 16000 * 3 + 77, 400, 0, 160 * 96 * 5 + 240, 16000 * 30])
def test_logmel_matches_oracle(engines, mel, n):
    e = engines("fp32")
    x = O.synth_audio(max(n, 1), seed=n % 97)[:n]
    xp = O.pad_waveform(x, 96)
    nf = 1 + (len(xp) - 400) // 160
    got = e.debug_logmel(x, nf)                       # virtual zero padding inside the kernel
    ref64 = O.log_mel(xp.astype(np.float64), mel, np.float64)
    ref32 = O.log_mel(xp, mel, np.float32)
    err_gpu = float(np.abs(got - ref64).max())
    err_o32 = float(np.abs(ref32 - ref64).max())
    _report(f"logmel_n{n}", {"gpu_vs_f64": err_gpu, "oracle32_vs_f64": err_o32,
                             "gpu_vs_oracle32": float(np.abs(got - ref32).max())})
    assert got.shape == ref32.shape
    assert np.isfinite(got).all()
    assert err_gpu <= max(2e-5, 4 * err_o32), (err_gpu, err_o32)


def test_logmel_unaligned_pointer_and_offsets(engines, mel):
    """frame_begin / tile boundaries: a long signal's frames equal those of its slices."""
    e = engines("fp32")
    x = O.synth_audio(16000 * 8, seed=5)
    nf = 1 + (len(x) - 400) // 160
    full = e.debug_logmel(x, nf)
    for f0 in (1, 31, 32, 33, 100):
        part = e.debug_logmel(x[f0 * 160:], nf - f0)
        assert np.array_equal(part, full[f0:]), f0


# ------------------------------------------------------------------------------------------------ pointwise GEMM
PW_SHAPES = [(1536 * 2, 64, 32), (384 * 3, 128, 64), (384, 128, 128), (96 * 5, 256, 128), (96 * 3 + 17, 256, 256),
             (24 * 11, 512, 256), (24 * 9 + 5, 512, 512), (6 * 40, 1024, 512), (6 * 33 + 1, 1024, 1024), (1, 64, 32),
             (128 * 150, 128, 128)]


@pytest.mark.parametrize("precision", ["fp32", "fp16x3", "fp16", "fp16f8"])
@pytest.mark.parametrize("M,N,K", PW_SHAPES)
def test_pw_gemm_matches_float64(engines, precision, M, N, K):
    if precision == "fp16f8" and K % 64:
        pytest.skip("the fp16 + fp8 plan works on 64-channel k-blocks")
    e = engines("fp32")
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = np.maximum(rng.standard_normal((M, K)), 0).astype(np.float32) * 2.0
    Wt = (rng.standard_normal((N, K)) * np.sqrt(2.0 / K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32) * 0.3
    ref = np.maximum(A.astype(np.float64) @ Wt.astype(np.float64).T + b, 0)
    scale = float(np.abs(ref).max())
    for bn in ([0] if precision == "fp32" else [0, 64] + ([128] if N % 128 == 0 else []) + ([256] if N % 256 == 0 else [])):
        got = e.debug_pw_gemm(A, Wt, b, precision, bn)
        err = float(np.abs(got - ref).max()) / scale
        _report(f"pw_{precision}_M{M}_N{N}_K{K}_bn{bn}", err)
        tol = {"fp32": 2e-6, "fp16x3": 4e-6, "fp16": 3e-3, "fp16f8": 2e-4}[precision]
        assert err <= tol, (precision, M, N, K, bn, err)


# ------------------------------------------------------------------------------------------------ layer by layer
@pytest.mark.parametrize("precision,fuse_mask", [("fp32", 0), ("fp16x3", 0), ("fp16", 0), ("fp16x3", 0x7FF),
                                                 ("fp16", 0x7FF), ("fp16x3", 0x10002), ("fp32", 0x10000),
                                                 ("fp16x3", 0x20002), ("fp16", 0x20000), ("fp16x3", 0x507FF),
                                                 ("fp16", 0x4003E), ("fp16x3", 0xC0006), ("fp16", 0x80000),
                                                 ("fp16f8", 0), ("fp16f8", 0x507FF), ("fp16f8", 0x1D07DE)])
def test_every_stage_matches_oracle(engines, yamnet_variables, mel, precision, fuse_mask):
    """fuse_mask 0: separate depthwise / pointwise kernels (every intermediate is observable);
    0x7FF: layers 2..12 run as ONE fused kernel each (depthwise outputs stay in shared memory)."""
    e = engines(precision, early_patches=8, late_patches=16, fuse_mask=fuse_mask)
    x = O.synth_audio(16000 * 5, seed=11)            # 6 patches
    taps = {}
    O.embed(x, yamnet_variables, mel, 96, taps=taps)
    worst = {}
    for stage in range(0, 28):
        fused_sep = stage >= 2 and stage % 2 == 0 and (fuse_mask >> (stage // 2 - 1)) & 1 and precision != "fp32"
        fused_c1 = stage == 1 and (fuse_mask & 0x10000) and not ((fuse_mask & 1) and precision != "fp32")
        fused_l12 = stage in (1, 2) and (fuse_mask & 0xA0000) and precision != "fp32"
        if fused_sep or fused_c1 or fused_l12:
            with pytest.raises(RuntimeError, match="fused away"):
                e.debug_stage(x, stage)
            continue
        if stage == 0:
            ref = taps["logmel"][:(6 - 1) * 96 + 96]
        elif stage == 1:
            ref = taps["L1"]
        else:
            L = stage // 2 + 1
            ref = taps[f"L{L}dw" if stage % 2 == 0 else f"L{L}"]
        got = e.debug_stage(x, stage)
        assert got.size == ref.size, (stage, got.size, ref.size)
        err = float(np.abs(got - ref.ravel()).max()) / max(float(np.abs(ref).max()), 1e-6)
        worst[stage] = err
    _report(f"stages_{precision}_fuse{fuse_mask:x}", worst)
    tol = {"fp16": 3e-2, "fp16f8": 1e-3}.get(precision, 1e-4)
    bad = {s: v for s, v in worst.items() if not v <= tol}
    assert not bad, bad


# ------------------------------------------------------------------------------------------------ whole path
def _median_threshold(a):
    return float(np.median(a[:, 8]))


@pytest.mark.parametrize("precision,hop,seconds,fuse_mask", [
    ("fp16x3", 96, 61.3, -1), ("fp16x3", 48, 61.3, -1), ("fp32", 96, 20.0, -1), ("fp16x3", 96, 0.5, -1),
    ("fp16x3", 96, 199.68, -1), ("fp16", 96, 61.3, -1), ("fp16x3", 96, 61.3, 0), ("fp16x3", 48, 33.1, 0x7FF),
    ("fp16x3", 48, 33.1, 0x10002), ("fp16", 96, 20.0, 0x20000), ("fp16x3", 96, 61.3, 0x5003E),
    ("fp16x3", 48, 33.1, 0x507FF), ("fp16x3", 96, 61.3, 0xC0006), ("fp16x3", 48, 33.1, 0xC0006),
    ("fp16x3", 96, 61.3, 0x10002), ("fp16x3", 96, 61.3, 0x1D07DE), ("fp16x3", 48, 33.1, 0x1D07FE),
    ("fp16", 96, 20.0, 0x1D07DE), ("fp16f8", 96, 61.3, -1), ("fp16f8", 48, 33.1, -1), ("fp16f8", 96, 61.3, 0),
    ("fp16f8", 96, 199.68, -1),
])
def test_predict_matches_oracle(engines, yamnet_variables, mel, head, precision, hop, seconds, fuse_mask):
    e = engines(precision, early_patches=16, late_patches=48, fuse_mask=fuse_mask)   # several early / late sub-batches
    n = int(round(seconds * 16000))
    x = O.synth_audio(n, seed=int(seconds * 10) + hop)
    act, emb = e.predict(x, hop, want_embeddings=True)
    want, wemb = O.predict(x, yamnet_variables, mel, head[0], head[1], hop, return_embeddings=True)
    assert act.shape == want.shape == (O.frame_counts(n, hop)[2], 13)
    a_err = float(np.abs(act - want).max())
    e_err = float(np.abs(emb - wemb).max() / np.abs(wemb).max())
    thr = _median_threshold(want)
    near = np.abs(want[:, 8] - thr) <= 1e-3
    flips = int(((act[:, 8] > thr) != (want[:, 8] > thr))[~near].sum())
    rounded_diff = int((np.round(act, 2) != np.round(want, 2)).sum())
    _report(f"predict_{precision}_hop{hop}_{seconds}s_fuse{fuse_mask}", {"act_max_abs": a_err, "emb_max_rel": e_err, "flips": flips,
                                                        "rounded_cells_differing": rounded_diff,
                                                        "cells": int(act.size), "act_range": [float(want.min()), float(want.max())]})
    if precision == "fp16":
        assert a_err <= 5e-3 * float(np.abs(want).max())
    else:
        assert a_err <= (2e-4 if precision == "fp16f8" else 1e-3), a_err      # fp16f8: the adoption bar of VERDICT r1 item 7
        assert e_err <= (4e-4 if precision == "fp16f8" else 1e-4), e_err
        assert flips == 0


def test_plugin_surface_matches_engine(engines, yamnet_variables, mel, head):
    """load_model -> ModelGeneralV3.predict(samples).numpy() (the call src/inference/worker.py:72 makes)."""
    from buzzdetect_b200.inference.models import load_model
    model = load_model("model_general_v3", framehop_prop=1, initialize=False)
    assert model.embedder.framehop_s == 0.96 and model.embedder.samplerate == 16000
    model.initialize()
    x = O.synth_audio(16000 * 12 + 5, seed=2)
    res = model.predict(x)
    a = res.numpy()
    want = O.predict(x, yamnet_variables, mel, head[0], head[1], 96)
    assert a.dtype == np.float32 and a.shape == want.shape
    assert float(np.abs(a - want).max()) <= 1e-3
    # torch CPU tensors (what a pinned-buffer streamer hands over) are accepted without a copy
    import torch
    a2 = model.predict(torch.from_numpy(x)).numpy()
    assert np.array_equal(a, a2)
    # half hop through the yamnet_k2 plugin (crashes in the reference, supported here)
    m2 = load_model("model_general_v3", framehop_prop=0.5, initialize=True)
    a3 = m2.predict(x).numpy()
    assert a3.shape[0] == O.frame_counts(len(x), 48)[2]
    k = min(len(a3[::2]), len(a))                                  # padding differs: whole-hop may add one frame
    assert float(np.abs(a3[::2][:k] - a[:k]).max()) <= 1e-3       # even half-hop frames are the whole-hop frames
    with pytest.raises(ValueError):
        load_model("model_general_v3", framehop_prop=0.3, initialize=True)


def test_chunk_independence_and_graph_replay(engines):
    """Per-chunk results do not depend on sub-batch sizes, CUDA-graph replay, or what ran before."""
    x = O.synth_audio(16000 * 40, seed=9)
    a = engines("fp16x3", early_patches=16, late_patches=48).predict(x, 96)
    b = engines("fp16x3", early_patches=8, late_patches=16).predict(x, 96)
    c = engines("fp16x3", early_patches=16, late_patches=48, use_graph=False).predict(x, 96)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    e = engines("fp16x3", early_patches=16, late_patches=48)
    y = O.synth_audio(16000 * 7, seed=10)
    e.predict(y, 96)
    assert np.array_equal(e.predict(x, 96), a)        # replayed graph, same bits


def test_async_slots(engines):
    """bd_submit_host / bd_wait with two chunks in flight give the same bits as the synchronous call."""
    import torch
    e = engines("fp16x3", early_patches=16, late_patches=48)
    xs = [torch.from_numpy(O.synth_audio(16000 * 10 + i * 1000, seed=20 + i)).pin_memory() for i in range(4)]
    want = [e.predict(x.numpy(), 96) for x in xs]
    outs = [torch.empty((w.shape[0], 13), dtype=torch.float32).pin_memory() for w in want]
    for i, x in enumerate(xs):
        slot = i % 2
        e.wait(slot)
        e.submit_ptr(slot, x.data_ptr(), x.numel(), 96, outs[i].data_ptr())
    e.wait(0)
    e.wait(1)
    for o, w in zip(outs, want):
        assert np.array_equal(o.numpy(), w)


def test_device_resident_entry_point(engines):
    import torch
    e = engines("fp16x3", early_patches=16, late_patches=48)
    x = O.synth_audio(16000 * 15, seed=4)
    want = e.predict(x, 96)
    dx = torch.from_numpy(x).cuda()
    dact = torch.empty((want.shape[0], 13), dtype=torch.float32, device="cuda")
    P = e.predict_device_ptr(dx.data_ptr(), dx.numel(), 96, dact.data_ptr())
    assert P == want.shape[0]
    assert np.array_equal(dact.cpu().numpy(), want)
    prof = e.profile_device_ptr(dx.data_ptr(), dx.numel(), 96)
    # defaults: layers 1+2 run as one kernel (counted as conv1), layers 3..6 and 8..12 run fused with their pointwise
    assert prof["pointwise"]["launches"] == 12 and prof["depthwise"]["launches"] == 3
    assert all(prof[k]["ms"] > 0 for k in ("frontend", "conv1", "depthwise", "pointwise", "pool_head"))
    assert [v["pw_launches"] for v in prof["layers"].values()] == [0] + [1] * 12
    assert [v["dw_launches"] for v in prof["layers"].values()] == [0] * 5 + [1] + [0] * 5 + [1, 1]


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_prefix_and_padding_properties(engines):
    """BASELINE configs[1] size (1 h, 3750 patches), default sub-batches: frame p depends only on samples
    [15360 p, 15360 p + 15600), so any prefix reproduces the same rows bit for bit, and explicit zero padding up to
    the padded length changes nothing."""
    e = engines("fp16x3")
    base = O.synth_audio(60 * 16000, seed=77)
    x = np.tile(base, 60)
    full = e.predict(x, 96)
    assert full.shape == (3750, 13) and np.isfinite(full).all()
    for k in (1, 100, 1025, 3000):
        part = e.predict(x[:15360 * k + 240], 96)
        assert part.shape[0] == k and np.array_equal(part, full[:k]), k
    ragged = x[:16000 * 33 + 1234]
    npad, _, P = O.frame_counts(len(ragged), 96)
    a = e.predict(ragged, 96)
    b = e.predict(np.concatenate([ragged, np.zeros(npad - len(ragged), np.float32)]), 96)
    assert a.shape[0] == P and np.array_equal(a, b)


def test_full_size_tensor_core_path_agrees_with_cuda_core_path(engines):
    """BASELINE configs[1] size: the default plan (fused tcgen05 kernels, fp16 hi/lo split) against the float32 CUDA-core
    mode of the same engine -- two disjoint sets of kernels -- on one hour of audio, default sub-batches."""
    x = np.tile(O.synth_audio(60 * 16000, seed=78), 60)
    a = engines("fp16x3").predict(x, 96)
    b = engines("fp32").predict(x, 96)
    assert a.shape == b.shape == (3750, 13)
    err = float(np.abs(a - b).max())
    _report("full_size_fp16x3_vs_fp32", {"max_abs": err, "cells_rounded_differently": int((np.round(a, 2) != np.round(b, 2)).sum())})
    assert err <= 1e-3, err
    thr = float(np.median(b[:, 8]))
    near = np.abs(b[:, 8] - thr) <= 1e-3
    assert int(((a[:, 8] > thr) != (b[:, 8] > thr))[~near].sum()) == 0


def test_day_long_single_call_has_no_index_overflow(engines):
    """configs[2] scale in ONE call: 24 h = 1.3824e9 samples (5.5 GB, byte offsets beyond 2**32).  The signal repeats
    every 15360 samples, so all 90000 frames but the last (which sees the zero padding) must be bit-identical."""
    import torch
    e = engines("fp16x3")
    period = torch.from_numpy(O.synth_audio(15360, seed=5)).cuda()
    n = 24 * 3600 * 16000
    x = period.repeat(n // 15360)
    assert x.numel() == n
    P = O.frame_counts(n, 96)[2]
    assert P == 90000
    act = torch.empty((P, 13), dtype=torch.float32, device="cuda")
    got = e.predict_device_ptr(x.data_ptr(), n, 96, act.data_ptr())
    assert got == P
    a = act.cpu().numpy()
    assert np.isfinite(a).all()
    assert (a[:-1] == a[0]).all()
    assert not np.array_equal(a[-1], a[0])          # the last frame ends in 240 samples of padding
    del x, act
    torch.cuda.empty_cache()
