"""CPU ORACLE (test infrastructure, NOT product code) for buzzdetect's inference hot path.

Restates, op for op, what the reference computes for one chunk of 16 kHz mono audio:

    pad -> STFT(400/160/512, periodic Hann) -> |.| -> mel[257,64] -> log(x+0.001) -> 96-frame patches
        -> MobileNet-v1 YAMNet (conv + 13 separable blocks, BN inference, ReLU) -> mean(H,W) -> Dense(13)

Reference anchors (paths relative to /root/reference):
  * pad_waveform ............ embedders/yamnet/features.py:82-108  (float32 div + ceil, as in the graph)
  * STFT / mel / log / patch  embedders/yamnet/features.py:22-79, params.py:24-51
  * layer stack ............. embedders/yamnet/yamnet.py:26-106 (_YAMNET_LAYER_DEFS :77-93), cut at
                              global_average_pooling2d per embedders/yamnet/BUILD.py:14-18
  * op-level spec ........... embedders/yamnet_k2/models/yamnet_{wholehop,halfhop}/saved_model.pb
                              (function __inference__wrapped_model_*)
  * head .................... models/model_general_v3/model.py:18-31 (MatMul + BiasAdd, linear)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product path (buzzdetect_b200/) never does.

PARITY STATUS: the reference ships no tests/golden vectors and TensorFlow cannot be installed here, so this
restatement is pinned against (a) the reference's own SavedModel graph executed node by node by
oracle/graph_exec.py (fixtures in tests/golden/), and (b) the data self-checks of SURVEY.md section 4.
Parity against the TensorFlow *runtime* is unpinned.
"""
from __future__ import annotations

import math

import numpy as np

SAMPLE_RATE = 16000
WIN = 400          # 25 ms
HOP = 160          # 10 ms
NFFT = 512
NBINS = 257
NMEL = 64
PATCH_FRAMES = 96
MIN_SAMPLES = 15600            # Const_5 in the graphs = int32(0.975 * 16000)
LOG_OFFSET = np.float32(0.001)
BN_EPS = 1e-4                  # params.py:48

# (kind, stride, cout) -- embedders/yamnet/yamnet.py:77-93
LAYER_DEFS = [("conv", 2, 32), ("sep", 1, 64), ("sep", 2, 128), ("sep", 1, 128), ("sep", 2, 256),
              ("sep", 1, 256), ("sep", 2, 512), ("sep", 1, 512), ("sep", 1, 512), ("sep", 1, 512),
              ("sep", 1, 512), ("sep", 1, 512), ("sep", 2, 1024), ("sep", 1, 1024)]


# ----------------------------------------------------------------------------------- framing math

def patch_hop_samples(hop_frames: int) -> int:
    """hop in samples as the graph holds it (Const_3): 15360 for wholehop, 7680 for halfhop."""
    return hop_frames * HOP


def pad_amount(n: int, hop_samples: int) -> int:
    """features.py:82-108.  The hop count uses a float32 division and ceil (graph: Cast/RealDiv/Ceil)."""
    pad = max(0, MIN_SAMPLES - n)
    n2 = max(n, MIN_SAMPLES)
    after = n2 - MIN_SAMPLES
    q = np.float32(after) / np.float32(hop_samples)
    hops = int(np.ceil(q))
    pad += hop_samples * hops - after
    return int(pad)


def frame_counts(n: int, hop_frames: int = PATCH_FRAMES):
    """(n_padded, n_stft_frames, n_patches) for n input samples; tf.signal.frame keeps complete frames only.

    pad can be negative when the float32 quotient rounds below the exact one (n - 15600 > 2**24): TF's Pad
    would then raise; we report it as is so callers can assert on it."""
    pad = pad_amount(n, patch_hop_samples(hop_frames))
    npad = n + pad
    f = 1 + (npad - WIN) // HOP if npad >= WIN else 0
    p = 1 + (f - PATCH_FRAMES) // hop_frames if f >= PATCH_FRAMES else 0
    return npad, f, p


def frame_starts(n_patches: int, framehop_s: float, time_start: float, digits_time: int = 2):
    """src/write/formatting.py:5-17: start = round(i*framehop_s + time_start, digits) on python floats
    (pandas float64 column; round-half-even as numpy.round)."""
    i = np.arange(n_patches, dtype=np.float64) * framehop_s
    if time_start != 0:
        i = i + time_start
    return np.round(i, digits_time)


# ----------------------------------------------------------------------------------- frontend

def hann_window(dtype=np.float32):
    """tf.signal.hann_window(400, periodic=True) as the graph computes it: 0.5 - 0.5*cos(2pi_f32*n/400)."""
    if dtype == np.float32:
        n = np.arange(WIN, dtype=np.float32)
        arg = (np.float32(6.2831855) * n) / np.float32(WIN)
        return (np.float32(0.5) - np.float32(0.5) * np.cos(arg, dtype=np.float32)).astype(np.float32)
    n = np.arange(WIN, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * n / WIN)


def pad_waveform(x: np.ndarray, hop_frames: int) -> np.ndarray:
    pad = pad_amount(len(x), patch_hop_samples(hop_frames))
    if pad < 0:
        raise ValueError("negative padding (float32 ceil quirk); TensorFlow's Pad would fail here too")
    return np.concatenate([x, np.zeros(pad, dtype=x.dtype)])


def log_mel(x_padded: np.ndarray, mel: np.ndarray, dtype=np.float32, block: int = 16384) -> np.ndarray:
    """[n_padded] -> [n_stft_frames, 64] log-mel.  fft zero-pads each 400-sample frame AT THE END to 512."""
    import scipy.fft
    x = np.ascontiguousarray(x_padded, dtype=dtype)
    n = len(x)
    nf = 1 + (n - WIN) // HOP
    win = hann_window(dtype)
    melm = mel.astype(dtype)
    out = np.empty((nf, NMEL), dtype=dtype)
    frames = np.lib.stride_tricks.as_strided(x, shape=(nf, WIN), strides=(HOP * x.itemsize, x.itemsize),
                                             writeable=False)
    off = dtype(0.001) if dtype == np.float64 else LOG_OFFSET
    for s in range(0, nf, block):
        fr = frames[s:s + block] * win
        spec = scipy.fft.rfft(fr, n=NFFT, axis=-1)           # float32 in -> complex64 out (pocketfft)
        mag = np.abs(spec).astype(dtype)
        out[s:s + block] = np.log(mag @ melm + off)
    return out


def patches_from_logmel(lm: np.ndarray, hop_frames: int) -> np.ndarray:
    nf = lm.shape[0]
    p = 1 + (nf - PATCH_FRAMES) // hop_frames if nf >= PATCH_FRAMES else 0
    idx = (np.arange(p)[:, None] * hop_frames + np.arange(PATCH_FRAMES)[None, :])
    return lm[idx]                                            # [P, 96, 64]


# ----------------------------------------------------------------------------------- MobileNet-v1 stack

def layer_tensor_names():
    """Checkpoint names per layer: returns list of dicts with keys dw, dw_bn, w, bn."""
    out = [{"w": "layer_with_weights-0/kernel", "bn": "layer_with_weights-1"}]
    for L in range(2, 15):
        b = 4 * (L - 2) + 2
        out.append({"dw": f"layer_with_weights-{b}/depthwise_kernel", "dw_bn": f"layer_with_weights-{b + 1}",
                    "w": f"layer_with_weights-{b + 2}/kernel", "bn": f"layer_with_weights-{b + 3}"})
    return out


def _same_pad(size: int, stride: int, k: int = 3):
    """TensorFlow SAME: total = max((ceil(size/stride)-1)*stride + k - size, 0); before = total//2."""
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


def _bn_relu(t, variables, prefix, torch):
    beta = torch.from_numpy(variables[prefix + "/beta"]).to(t.dtype)
    mean = torch.from_numpy(variables[prefix + "/moving_mean"]).to(t.dtype)
    var = torch.from_numpy(variables[prefix + "/moving_variance"]).to(t.dtype)
    # FusedBatchNormV3 inference, scale == 1: y = (x - mean) * rsqrt(var + eps) + beta
    inv = torch.rsqrt(var + torch.tensor(BN_EPS, dtype=t.dtype))
    y = (t - mean.view(1, -1, 1, 1)) * inv.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)
    return torch.relu(y)


def mobilenet_embed(patches: np.ndarray, variables: dict, dtype=np.float32, batch: int = 64,
                    taps: dict | None = None) -> np.ndarray:
    """[P,96,64] log-mel patches -> [P,1024] embeddings.  torch-CPU convolutions (NCHW internally; the
    arithmetic is layout independent).  `taps`, if given, receives per-layer NHWC outputs of the first batch."""
    import torch
    import torch.nn.functional as F
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    names = layer_tensor_names()
    outs = []
    with torch.no_grad():
        for s in range(0, patches.shape[0], batch):
            t = torch.from_numpy(np.ascontiguousarray(patches[s:s + batch])).to(tdt).unsqueeze(1)  # [B,1,96,64]
            for li, ((kind, stride, cout), nm) in enumerate(zip(LAYER_DEFS, names)):
                H, W = t.shape[2], t.shape[3]
                ph, pw_ = _same_pad(H, stride), _same_pad(W, stride)
                if kind == "conv":
                    w = torch.from_numpy(variables[nm["w"]]).to(tdt).permute(3, 2, 0, 1)       # HWIO -> OIHW
                    t = F.conv2d(F.pad(t, (pw_[0], pw_[1], ph[0], ph[1])), w, stride=stride)
                    t = _bn_relu(t, variables, nm["bn"], torch)
                    if taps is not None and s == 0:
                        taps[f"L{li + 1}"] = t.permute(0, 2, 3, 1).numpy().copy()
                else:
                    C = t.shape[1]
                    dw = torch.from_numpy(variables[nm["dw"]]).to(tdt).permute(2, 3, 0, 1)     # [3,3,C,1] -> [C,1,3,3]
                    t = F.conv2d(F.pad(t, (pw_[0], pw_[1], ph[0], ph[1])), dw, stride=stride, groups=C)
                    t = _bn_relu(t, variables, nm["dw_bn"], torch)
                    if taps is not None and s == 0:
                        taps[f"L{li + 1}dw"] = t.permute(0, 2, 3, 1).numpy().copy()
                    w = torch.from_numpy(variables[nm["w"]]).to(tdt).permute(3, 2, 0, 1)       # [1,1,Ci,Co] -> [Co,Ci,1,1]
                    t = F.conv2d(t, w)
                    t = _bn_relu(t, variables, nm["bn"], torch)
                    if taps is not None and s == 0:
                        taps[f"L{li + 1}"] = t.permute(0, 2, 3, 1).numpy().copy()
            outs.append(t.mean(dim=(2, 3)).numpy())
    if not outs:
        return np.zeros((0, 1024), dtype=dtype)
    return np.concatenate(outs, axis=0)


def head(emb: np.ndarray, kernel: np.ndarray, bias: np.ndarray, dtype=np.float32) -> np.ndarray:
    return emb.astype(dtype) @ kernel.astype(dtype) + bias.astype(dtype)


# ----------------------------------------------------------------------------------- whole path

def embed(samples: np.ndarray, variables: dict, mel: np.ndarray, hop_frames: int = PATCH_FRAMES,
          dtype=np.float32, taps: dict | None = None):
    """YamnetK2.embed / EmbedderYamnet.embed (embedders/yamnet_k2/embedder.py:27-37): f32[n] -> [P,1024]."""
    x = pad_waveform(np.asarray(samples, dtype=np.float32), hop_frames)
    lm = log_mel(x, mel, dtype)
    if taps is not None:
        taps["logmel"] = lm
    p = patches_from_logmel(lm, hop_frames)
    return mobilenet_embed(p, variables, dtype, taps=taps)


def predict(samples: np.ndarray, variables: dict, mel: np.ndarray, head_kernel: np.ndarray,
            head_bias: np.ndarray, hop_frames: int = PATCH_FRAMES, dtype=np.float32, return_embeddings=False):
    """ModelGeneralV3.predict (models/model_general_v3/model.py:18-31): f32[n] -> [P,13] raw activations."""
    e = embed(samples, variables, mel, hop_frames, dtype)
    a = head(e, head_kernel, head_bias, dtype)
    return (a, e) if return_embeddings else a


# ----------------------------------------------------------------------------------- writer semantics

def format_activations(results: np.ndarray, digits_results: int = 2) -> np.ndarray:
    """src/write/formatting.py:30-31: np.array(results).round(digits) on float32 values."""
    return np.array(results).round(digits_results)


def format_detections(results: np.ndarray, threshold: float, buzz_index: int = 8) -> np.ndarray:
    """src/write/formatting.py:20-24: strict '>' on the raw (unrounded) float32 activation."""
    return (results[:, buzz_index] > threshold).astype(int)


def synth_audio(n: int, seed: int = 0, sr: int = SAMPLE_RATE) -> np.ndarray:
    """pink-ish noise (sigma 0.05) + three harmonic 'buzz' bursts, float32."""
    rng = np.random.default_rng(seed)
    white = rng.standard_normal(n).astype(np.float32)
    # one-pole low-pass mix ~ pink-ish tilt, cheap and deterministic
    from scipy.signal import lfilter
    pink = lfilter([0.05], [1.0, -0.95], white).astype(np.float32)
    x = 0.03 * white + pink * (0.05 / max(float(pink.std()), 1e-9))
    t = np.arange(n, dtype=np.float64) / sr
    dur = n / sr
    for k in range(3):
        f0 = 200.0 + 50.0 * k
        c = dur * (k + 1) / 4.0
        w = max(dur / 12.0, 0.3)
        env = np.exp(-0.5 * ((t - c) / w) ** 2)
        burst = sum(np.sin(2 * np.pi * f0 * h * t) / h for h in range(1, 6))
        x = x + (0.1 * env * burst).astype(np.float32)
    return np.clip(x, -1.0, 1.0).astype(np.float32)
