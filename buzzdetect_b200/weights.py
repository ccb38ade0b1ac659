"""Weight handling for the YAMNet embedder and the model_general_v3 head.

What the reference ships: the head weights are complete; the YAMNet
``variables.data-00000-of-00001`` blob is NOT in the checkout (``.MISSING_LARGE_BLOBS``), only its index.
This module therefore

* loads a real YAMNet blob when one is available (``$BUZZ_YAMNET_WEIGHTS`` or the reference's own relative
  paths), checking every tensor against the masked CRC32C recorded in ``variables.index``
  (committed as ``assets/yamnet_tensor_table.json``);
* otherwise builds clearly-labelled, deterministic SYNTHETIC weights of the exact shapes (seeded numpy
  generator, BN statistics from a committed calibration table) so that GPU-vs-oracle parity and throughput
  can still be measured;
* folds inference BatchNorm into the preceding convolution (w' = w*rsqrt(var+1e-4),
  b' = beta - mean*rsqrt(var+1e-4); scale is absent, embedders/yamnet/params.py:46-48) and lays the result
  out the way the CUDA kernels want it.
"""
from __future__ import annotations

import json
import os
import struct
from dataclasses import dataclass

import numpy as np

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")
BN_EPS = 1e-4

# (kind, stride, cin, cout, H_in, W_in) -- embedders/yamnet/yamnet.py:77-93 on a [96,64,1] patch
LAYERS = []
_h, _w, _c = 96, 64, 1
for _kind, _s, _co in [("conv", 2, 32), ("sep", 1, 64), ("sep", 2, 128), ("sep", 1, 128), ("sep", 2, 256),
                       ("sep", 1, 256), ("sep", 2, 512), ("sep", 1, 512), ("sep", 1, 512), ("sep", 1, 512),
                       ("sep", 1, 512), ("sep", 1, 512), ("sep", 2, 1024), ("sep", 1, 1024)]:
    LAYERS.append((_kind, _s, _c, _co, _h, _w))
    _h, _w, _c = -(-_h // _s), -(-_w // _s), _co
del _h, _w, _c, _kind, _s, _co


def tensor_table() -> dict:
    with open(os.path.join(ASSETS, "yamnet_tensor_table.json")) as f:
        return json.load(f)


def layer_tensor_names():
    """Checkpoint tensor names per layer (object-graph order)."""
    out = [{"w": "layer_with_weights-0/kernel", "bn": "layer_with_weights-1"}]
    for L in range(2, 15):
        b = 4 * (L - 2) + 2
        out.append({"dw": f"layer_with_weights-{b}/depthwise_kernel", "dw_bn": f"layer_with_weights-{b + 1}",
                    "w": f"layer_with_weights-{b + 2}/kernel", "bn": f"layer_with_weights-{b + 3}"})
    return out


# ----------------------------------------------------------------------------------- crc32c

_CRC_TBL = None


def _crc32c(data: bytes) -> int:
    global _CRC_TBL
    if _CRC_TBL is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TBL = t
    t = _CRC_TBL
    c = 0xFFFFFFFF
    for b in data:
        c = t[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    """CRC as stored in BundleEntryProto.crc32c."""
    c = _crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------------- real weights

_REL_BLOBS = (
    "embedders/yamnet_k2/models/yamnet_wholehop/variables/variables.data-00000-of-00001",
    "embedders/yamnet_k2/models/yamnet_halfhop/variables/variables.data-00000-of-00001",
    "embedders/yamnet/variables/variables.data-00000-of-00001",
)


def find_yamnet_blob() -> str | None:
    """$BUZZ_YAMNET_WEIGHTS, then the reference's own relative locations under CWD and $BUZZDETECT_ROOT."""
    env = os.environ.get("BUZZ_YAMNET_WEIGHTS")
    if env:
        if not os.path.exists(env):
            raise FileNotFoundError(f"BUZZ_YAMNET_WEIGHTS={env} does not exist")
        return env
    roots = [os.getcwd()]
    if os.environ.get("BUZZDETECT_ROOT"):
        roots.append(os.environ["BUZZDETECT_ROOT"])
    for root in roots:
        for rel in _REL_BLOBS:
            p = os.path.join(root, rel)
            if os.path.exists(p):
                return p
    return None


def load_yamnet_blob(path: str, verify: bool = True) -> dict:
    """Read a TF tensor-bundle data shard using the committed tensor table; CRC-check each tensor."""
    tab = tensor_table()
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < tab["float_bytes"]:
        raise ValueError(f"{path}: {len(buf)} bytes, expected at least {tab['float_bytes']}")
    out = {}
    for t in tab["tensors"]:
        blob = buf[t["offset"]:t["offset"] + t["size"]]
        if verify and masked_crc32c(blob) != t["crc32c"]:
            raise ValueError(f"{path}: CRC32C mismatch for tensor {t['name']}")
        out[t["name"]] = np.frombuffer(blob, dtype="<f4").reshape(t["shape"]).copy()
    return out


# ----------------------------------------------------------------------------------- synthetic weights

# Per-layer (input mean after ReLU, pre-BN variance after centring) measured once by
# tools/calibrate_synthetic.py on oracle.synth_audio(16000*20, seed=123) with SYNTH_SEED below.  They only
# keep the random network's activations O(1) from layer to layer, like a freshly BN-calibrated net.
SYNTH_SEED = 20251018
SYNTH_EMB_SCALE = 0.125
SYNTH_CALIB: list[tuple[float, float]] = []   # filled below from assets/synth_calibration.json if present
_calib_path = os.path.join(ASSETS, "synth_calibration.json")
if os.path.exists(_calib_path):
    with open(_calib_path) as _f:
        SYNTH_CALIB = [tuple(x) for x in json.load(_f)["stages"]]


def synth_stage_names():
    """The 27 conv stages in execution order: ('L1','conv'), ('L2','dw'), ('L2','pw'), ..."""
    out = [("L1", "conv")]
    for L in range(2, 15):
        out += [(f"L{L}", "dw"), (f"L{L}", "pw")]
    return out


def synthetic_yamnet(seed: int = SYNTH_SEED, calib: list | None = None, upto: int | None = None) -> dict:
    """Deterministic stand-in for the missing YAMNet blob.  SAME names and shapes as the real checkpoint.

    kernels ~ N(0, 2/fan_in); moving_mean[c] = mu_in * sum(kernel[..., c]); moving_variance[c] = V * U(0.5,2);
    beta[c] ~ N(0.1, 0.3).  (mu_in, V) per stage come from the calibration table."""
    calib = SYNTH_CALIB if calib is None else calib
    rng = np.random.default_rng(seed)
    names = layer_tensor_names()
    out = {}
    stage = 0

    def bn(prefix, w_sum, c):
        nonlocal stage
        mu_in, v = calib[stage] if stage < len(calib) else (0.0, 1.0)
        beta = 0.1 + 0.3 * rng.standard_normal(c)
        var = v * rng.uniform(0.5, 2.0, c)
        if stage == 26:
            # last stage: shrink the embeddings so that the REAL model_general_v3 head produces logits in the range
            # the reference documents (about -6 .. +1.5, models/model_general_v3/tests/metrics.csv); the absolute
            # 1e-3 activation tolerance is then tested at a realistic scale.
            beta = beta * SYNTH_EMB_SCALE
            var = var / SYNTH_EMB_SCALE ** 2
        out[prefix + "/beta"] = beta.astype(np.float32)
        out[prefix + "/moving_mean"] = (mu_in * w_sum).astype(np.float32)
        out[prefix + "/moving_variance"] = var.astype(np.float32)
        stage += 1

    for (kind, s, cin, cout, H, W), nm in zip(LAYERS, names):
        if upto is not None and stage >= upto:
            break
        if kind == "conv":
            w = (rng.standard_normal((3, 3, cin, cout)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32)
            out[nm["w"]] = w
            bn(nm["bn"], w.sum(axis=(0, 1, 2), dtype=np.float64), cout)
        else:
            dw = (rng.standard_normal((3, 3, cin, 1)) * np.sqrt(2.0 / 9)).astype(np.float32)
            out[nm["dw"]] = dw
            bn(nm["dw_bn"], dw.sum(axis=(0, 1, 3), dtype=np.float64), cin)
            if upto is not None and stage >= upto:
                break
            w = (rng.standard_normal((1, 1, cin, cout)) * np.sqrt(2.0 / cin)).astype(np.float32)
            out[nm["w"]] = w
            bn(nm["bn"], w.sum(axis=(0, 1, 2), dtype=np.float64), cout)
    return out


# ----------------------------------------------------------------------------------- head + mel

def load_head():
    k = np.fromfile(os.path.join(ASSETS, "head_kernel_1024x13.f32"), dtype="<f4").reshape(1024, 13)
    b = np.fromfile(os.path.join(ASSETS, "head_bias_13.f32"), dtype="<f4")
    return k, b


def load_mel(embedder: str = "yamnet_k2") -> np.ndarray:
    """The mel matrix the graphs carry as Const_1 (do NOT recompute: SURVEY.md section 4 (ii))."""
    fn = "mel_yamnet_257x64.f32" if embedder == "yamnet" else "mel_257x64.f32"
    return np.fromfile(os.path.join(ASSETS, fn), dtype="<f4").reshape(257, 64)


# ----------------------------------------------------------------------------------- BN folding / packing

@dataclass
class FoldedLayer:
    kind: str            # "conv" | "sep"
    stride: int
    cin: int
    cout: int
    h_in: int
    w_in: int
    dw_w: np.ndarray | None      # [9, cin]  (tap-major, channel-contiguous)
    dw_b: np.ndarray | None      # [cin]
    w: np.ndarray                # conv: [9, cout]; sep: [cout, cin] (K-major B operand for the GEMM)
    b: np.ndarray                # [cout]


def _fold(kernel64: np.ndarray, bn: dict, axis_out: int):
    inv = 1.0 / np.sqrt(bn["moving_variance"].astype(np.float64) + BN_EPS)
    shape = [1] * kernel64.ndim
    shape[axis_out] = -1
    w = kernel64 * inv.reshape(shape)
    b = bn["beta"].astype(np.float64) - bn["moving_mean"].astype(np.float64) * inv
    return w, b


def fold_yamnet(variables: dict) -> list[FoldedLayer]:
    """Fold BN in float64, round once to float32."""
    out = []
    for (kind, s, cin, cout, H, W), nm in zip(LAYERS, layer_tensor_names()):
        def bn(prefix):
            return {k: variables[f"{prefix}/{k}"] for k in ("beta", "moving_mean", "moving_variance")}
        if kind == "conv":
            w, b = _fold(variables[nm["w"]].astype(np.float64), bn(nm["bn"]), 3)       # [3,3,1,32]
            out.append(FoldedLayer(kind, s, cin, cout, H, W, None, None,
                                   np.ascontiguousarray(w.reshape(9, cout), dtype=np.float32),
                                   b.astype(np.float32)))
        else:
            dw, dwb = _fold(variables[nm["dw"]].astype(np.float64), bn(nm["dw_bn"]), 2)  # [3,3,C,1]
            w, b = _fold(variables[nm["w"]].astype(np.float64), bn(nm["bn"]), 3)         # [1,1,Ci,Co]
            out.append(FoldedLayer(kind, s, cin, cout, H, W,
                                   np.ascontiguousarray(dw.reshape(9, cin), dtype=np.float32),
                                   dwb.astype(np.float32),
                                   np.ascontiguousarray(w.reshape(cin, cout).T, dtype=np.float32),
                                   b.astype(np.float32)))
    return out


def pack_folded(layers: list[FoldedLayer]) -> tuple[np.ndarray, list[dict]]:
    """One flat float32 blob + offsets (in floats) in the order the C ABI expects (include/buzzdetect_b200.h)."""
    chunks, meta, off = [], [], 0

    def add(a):
        nonlocal off
        a = np.ascontiguousarray(a, dtype=np.float32).ravel()
        pad = (-a.size) % 64                         # keep every tensor 256-byte aligned
        chunks.append(a)
        if pad:
            chunks.append(np.zeros(pad, dtype=np.float32))
        o = off
        off += a.size + pad
        return o

    for l in layers:
        m = {"kind": l.kind, "stride": l.stride, "cin": l.cin, "cout": l.cout, "h_in": l.h_in, "w_in": l.w_in}
        if l.kind == "sep":
            m["dw_w"] = add(l.dw_w)
            m["dw_b"] = add(l.dw_b)
        m["w"] = add(l.w)
        m["b"] = add(l.b)
        meta.append(m)
    return np.concatenate(chunks), meta


def synthetic_allowed(flag: bool | None = None) -> bool:
    """Synthetic YAMNet weights are an explicit opt-in: tests, bench.py and smoke() pass allow_synthetic=True or set
    BUZZ_B200_ALLOW_SYNTHETIC=1.  A production run without the real blob must fail, not write detections from a random
    network."""
    if flag is not None:
        return bool(flag)
    return os.environ.get("BUZZ_B200_ALLOW_SYNTHETIC", "") not in ("", "0")


def resolve_yamnet(verify: bool = True, allow_synthetic: bool | None = None) -> tuple[dict, str]:
    """(variables, provenance) -- provenance is 'real:<path>' or 'synthetic:<seed>'."""
    p = find_yamnet_blob()
    if p is not None:
        return load_yamnet_blob(p, verify=verify), f"real:{p}"
    if not synthetic_allowed(allow_synthetic):
        raise FileNotFoundError(
            "YAMNet weights not found: set BUZZ_YAMNET_WEIGHTS to variables.data-00000-of-00001 (or run from a "
            "buzzdetect checkout / set BUZZDETECT_ROOT).  The checkout this package was developed against ships only "
            "variables.index (.MISSING_LARGE_BLOBS).  Seeded synthetic weights exist for tests and benchmarks only: "
            "Engine(allow_synthetic=True) or BUZZ_B200_ALLOW_SYNTHETIC=1.")
    return synthetic_yamnet(), f"synthetic:{SYNTH_SEED}"
