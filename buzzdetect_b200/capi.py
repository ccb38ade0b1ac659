"""ctypes binding of libbuzzdetect_b200.so (include/buzzdetect_b200.h).

This is the thin layer the north star asks for: Python host -> C ABI -> hand-written CUDA.  There is NO CPU
fallback: if the shared library is missing, or no sm_100 device is present, construction raises.
"""
from __future__ import annotations

import collections
import ctypes as C
import os
import threading
import weakref

import numpy as np

from . import weights as W

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbuzzdetect_b200.so")

PRECISION = {"fp32": 0, "fp32_simt": 0, "fp16": 1, "fp16x1": 1, "fp16f8": 2, "fp16x3": 3}
DEFAULT_PRECISION = os.environ.get("BUZZ_B200_PRECISION", "fp16x3")

N_LAYERS = 14
EMBED_DIM = 1024
STAGES = ("frontend", "conv1", "depthwise", "pointwise", "pool_head")


class bd_layer_desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("stride", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
                ("h_in", C.c_int32), ("w_in", C.c_int32), ("dw_w", C.c_int64), ("dw_b", C.c_int64),
                ("w", C.c_int64), ("b", C.c_int64)]


class bd_weights(C.Structure):
    _fields_ = [("folded", C.POINTER(C.c_float)), ("folded_len", C.c_int64), ("layers", bd_layer_desc * N_LAYERS),
                ("mel", C.POINTER(C.c_float)), ("window", C.POINTER(C.c_float)),
                ("head_kernel", C.POINTER(C.c_float)), ("head_bias", C.POINTER(C.c_float)),
                ("n_classes", C.c_int32)]


class bd_config(C.Structure):
    _fields_ = [("device", C.c_int32), ("precision", C.c_int32), ("early_patches", C.c_int32),
                ("late_patches", C.c_int32), ("use_graph", C.c_int32), ("n_slots", C.c_int32),
                ("fuse_mask", C.c_int32)]


_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); also the list tests/test_abi.py checks against the header
SIGNATURES = {
    "bd_abi_version": (C.c_int32, []),
    "bd_frames_for": (C.c_int32, [C.c_int64, C.c_int32, _i64p, _i64p, _i64p]),
    "bd_engine_create": (C.c_int32, [C.POINTER(bd_config), C.POINTER(bd_weights), C.POINTER(C.c_void_p), C.c_char_p,
                                     C.c_size_t]),
    "bd_engine_destroy": (None, [C.c_void_p]),
    "bd_last_error": (C.c_char_p, [C.c_void_p]),
    "bd_predict_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, _i64p]),
    "bd_predict_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, _i64p]),
    "bd_submit_host": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                   _i64p]),
    "bd_submit_pcm_host": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                       C.c_int32, C.c_void_p, C.c_void_p, _i64p]),
    "bd_wait": (C.c_int32, [C.c_void_p, C.c_int32]),
    "bd_flush": (C.c_int32, [C.c_void_p]),
    "bd_reserve_slots": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int64, C.c_int32]),
    "bd_debug_stats": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "bd_trace": (C.c_int32, [C.c_void_p, C.c_int32, C.c_char_p, C.c_size_t]),
    "bd_set_auto_flush": (C.c_int32, [C.c_void_p, C.c_int32]),
    "bd_slot_state": (C.c_int32, [C.c_void_p, C.c_int32]),
    "bd_batch_stats": (C.c_int32, [C.c_void_p, _i64p, _i64p]),
    "bd_host_alloc": (C.c_int32, [C.c_size_t, C.c_int32, C.POINTER(C.c_void_p)]),
    "bd_host_free": (None, [C.c_void_p]),
    "bd_synchronize": (C.c_int32, [C.c_void_p]),
    "bd_profile_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, _f32p, _i64p]),
    "bd_launch_count": (C.c_int64, [C.c_void_p]),
    "bd_bench_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, _f32p]),
    "bd_resample_out_len": (C.c_int64, [C.c_int64, C.c_int32]),
    "bd_resample_host": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p,
                                     C.c_int64, _i64p]),
    "bd_resample_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                       C.c_void_p, C.c_int64, _i64p]),
    "bd_debug_logmel": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "bd_debug_pw_gemm": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_void_p]),
    "bd_debug_stage": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                   _i64p]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library(path: str | None = None):
    """dlopen the CUDA library; raises if it has not been built (no fallback of any kind)."""
    global _lib
    with _lib_lock:
        if _lib is None:
            p = path or os.environ.get("BUZZ_B200_LIB", LIB_PATH)
            if not os.path.exists(p):
                raise RuntimeError(
                    f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a). buzzdetect_b200 has no CPU fallback.")
            lib = C.CDLL(p)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            if lib.bd_abi_version() != 1:
                raise RuntimeError("libbuzzdetect_b200.so ABI version mismatch")
            _lib = lib
    return _lib


def frames_for(n_samples: int, hop_frames: int = 96):
    """(n_padded, n_stft_frames, n_patches) -- host integer / float32-ceil math, no GPU needed."""
    lib = load_library()
    a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
    if lib.bd_frames_for(int(n_samples), int(hop_frames), C.byref(a), C.byref(b), C.byref(c)):
        raise ValueError(f"bd_frames_for({n_samples}, {hop_frames}) rejected its arguments")
    return a.value, b.value, c.value


def hop_frames_for(framehop_prop: float) -> int:
    """Patch hop in STFT frames exactly as features.py:70-71 computes it: int(round(100 * patch_hop_seconds))."""
    hop_s = 0.96 * framehop_prop
    return int(round((16000.0 / 160) * hop_s))


def hann_window_f32() -> np.ndarray:
    """tf.signal.hann_window(400, periodic=True) with the graph's float32 op sequence."""
    n = np.arange(400, dtype=np.float32)
    arg = (np.float32(6.2831855) * n) / np.float32(400)
    return (np.float32(0.5) - np.float32(0.5) * np.cos(arg, dtype=np.float32)).astype(np.float32)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def pinned_empty(shape, dtype=np.float32, write_combined: bool = False) -> np.ndarray:
    """numpy array over pinned (page-locked) host memory from bd_host_alloc: what the streamer's chunk ring is made of.
    The memory is released when the last view of the array is garbage collected.  Needs a CUDA device."""
    lib = load_library()
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(d) for d in shape)
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
    p = C.c_void_p()
    if lib.bd_host_alloc(max(nbytes, 1), 1 if write_combined else 0, C.byref(p)) or not p.value:
        msg = lib.bd_last_error(None)
        raise RuntimeError("bd_host_alloc: " + (msg.decode(errors="replace") if msg else "failed"))
    buf = (C.c_char * max(nbytes, 1)).from_address(p.value)
    weakref.finalize(buf, lib.bd_host_free, p.value)
    return np.frombuffer(buf, dtype=dt, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)


class Ticket:
    """One chunk in flight (bd_submit_* ... bd_wait).  result() blocks until the activations are on the host; until
    then the ticket keeps the input buffer alive (the DMA engine may still be reading it)."""

    __slots__ = ("_eng", "slot", "act", "emb", "n_samples", "_keep", "_done", "_lock", "__weakref__")

    def __init__(self, eng, slot, act, emb, n_samples, keep):
        self._eng, self.slot, self.act, self.emb, self.n_samples, self._keep = eng, slot, act, emb, n_samples, keep
        self._done = False
        self._lock = threading.Lock()

    def done(self) -> bool:
        return self._done

    def wait(self):
        with self._lock:
            if not self._done:
                try:
                    self._eng._wait_slot(self.slot)
                finally:
                    self._done = True
                    self._keep = None
                    self._eng._release(self)
        return self

    def result(self):
        self.wait()
        return self.act if self.emb is None else (self.act, self.emb)


class Engine:
    """One inferer: YAMNet frontend + MobileNet-v1 + dense head resident on one B200."""

    def __init__(self, device: int = 0, yamnet_variables: dict | None = None, embedder: str = "yamnet_k2",
                 head: tuple | None = None, precision: str | None = None, early_patches: int = 0,
                 late_patches: int = 0, use_graph: bool = True, n_slots: int = 2, fuse_mask: int = -1,
                 allow_synthetic: bool | None = None, max_inflight_samples: int = 1 << 27):
        self._h = None
        lib = load_library()
        self._lib = lib
        precision = precision or DEFAULT_PRECISION
        if precision not in PRECISION:
            raise ValueError(f"precision must be one of {sorted(PRECISION)}")
        self.precision = precision
        if yamnet_variables is None:
            yamnet_variables, self.weights_provenance = W.resolve_yamnet(allow_synthetic=allow_synthetic)
        else:
            self.weights_provenance = "caller"
        folded = W.fold_yamnet(yamnet_variables)
        blob, meta = W.pack_folded(folded)
        hk, hb = head if head is not None else W.load_head()
        hk = np.ascontiguousarray(hk, dtype=np.float32)
        hb = np.ascontiguousarray(hb, dtype=np.float32)
        mel = np.ascontiguousarray(W.load_mel(embedder), dtype=np.float32)
        win = hann_window_f32()
        self.n_classes = int(hb.shape[0])
        self.device = int(device)

        w = bd_weights()
        w.folded = blob.ctypes.data_as(_f32p)
        w.folded_len = blob.size
        for i, m in enumerate(meta):
            d = w.layers[i]
            d.kind = 0 if m["kind"] == "conv" else 1
            d.stride, d.cin, d.cout, d.h_in, d.w_in = m["stride"], m["cin"], m["cout"], m["h_in"], m["w_in"]
            d.dw_w, d.dw_b = m.get("dw_w", -1), m.get("dw_b", -1)
            d.w, d.b = m["w"], m["b"]
        w.mel = mel.ctypes.data_as(_f32p)
        w.window = win.ctypes.data_as(_f32p)
        w.head_kernel = hk.ctypes.data_as(_f32p)
        w.head_bias = hb.ctypes.data_as(_f32p)
        w.n_classes = self.n_classes
        cfg = bd_config(self.device, PRECISION[precision], early_patches, late_patches, 1 if use_graph else 0, n_slots,
                        fuse_mask)
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        if lib.bd_engine_create(C.byref(cfg), C.byref(w), C.byref(h), err, 512):
            raise RuntimeError("bd_engine_create: " + err.value.decode(errors="replace"))
        self._h = h
        self.n_slots = max(1, min(int(n_slots) if n_slots > 0 else 2, 64))
        # ticket API (submit / submit_pcm): free slots, tickets in submission order, in-flight budget
        self._tk_lock = threading.Lock()
        self._free = list(range(self.n_slots - 1, -1, -1))
        self._outstanding = collections.deque()
        self._inflight_samples = 0
        self.max_inflight_samples = int(max_inflight_samples)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc, what):
        if rc:
            msg = self._lib.bd_last_error(self._h)
            raise RuntimeError(f"{what}: {msg.decode(errors='replace') if msg else 'unknown error'}")

    def close(self):
        if getattr(self, "_h", None):
            try:
                self.drain()
            except Exception:
                pass
            self._lib.bd_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self._lib.bd_launch_count(self._h))

    # ------------------------------------------------------------------ hot path
    def predict(self, samples: np.ndarray, hop_frames: int = 96, want_embeddings: bool = False):
        """Host numpy in, host numpy out: [P, n_classes] (and [P,1024])."""
        return self.submit(samples, hop_frames, want_embeddings).result()

    # ------------------------------------------------------------------ tickets: chunks in flight
    def _wait_slot(self, slot: int):
        self._check(self._lib.bd_wait(self._h, slot), "bd_wait")

    def _release(self, tk: "Ticket"):
        with self._tk_lock:
            try:
                self._outstanding.remove(tk)
            except ValueError:
                pass
            self._inflight_samples -= tk.n_samples
            self._free.append(tk.slot)

    def _acquire_slot(self, n_samples: int) -> int:
        """A free slot; completes the oldest tickets first when every slot is taken or the in-flight budget is spent
        (back-pressure on the submitting thread, like the reference's bounded q_analyze)."""
        while True:
            with self._tk_lock:
                over = self._outstanding and self._inflight_samples + n_samples > self.max_inflight_samples
                if self._free and not over:
                    self._inflight_samples += n_samples
                    return self._free.pop()
                oldest = self._outstanding[0] if self._outstanding else None
            if oldest is None:
                raise RuntimeError("no free slot: slots are held through the raw submit_ptr/wait API")
            oldest.wait()

    def _register(self, slot, act, emb, n_samples, keep) -> "Ticket":
        tk = Ticket(self, slot, act, emb, n_samples, keep)
        with self._tk_lock:
            self._outstanding.append(tk)
        return tk

    def submit(self, samples: np.ndarray, hop_frames: int = 96, want_embeddings: bool = False) -> "Ticket":
        """Queue one 16 kHz mono float32 chunk; returns at once (pinned input) with a Ticket.  Chunks queued while the
        GPU is busy run together as one pass (include/buzzdetect_b200.h, "Chunk coalescing")."""
        x = np.ascontiguousarray(samples, dtype=np.float32)
        if x.ndim != 1:
            raise ValueError("samples must be 1-D (mono, 16 kHz)")
        _, _, P = frames_for(x.size, hop_frames)
        act = np.empty((P, self.n_classes), dtype=np.float32)
        emb = np.empty((P, EMBED_DIM), dtype=np.float32) if want_embeddings else None
        if x.size > getattr(self, "_reserved", 0) and not self._outstanding:
            self.reserve(x.size, getattr(self, "_reserved_pcm", 0), hop_frames)
        slot = self._acquire_slot(x.size)
        tk = self._register(slot, act, emb, x.size, x)
        try:
            got = self.submit_ptr(slot, x.ctypes.data, x.size, hop_frames, act.ctypes.data,
                                  emb.ctypes.data if emb is not None else None)
        except Exception:
            tk._done = True
            self._release(tk)
            raise
        assert got == P
        return tk

    def submit_pcm(self, pcm: np.ndarray, src_rate: int, hop_frames: int = 96) -> "Ticket":
        """Queue one chunk still in its decoded form: [n] or [n, channels], int16 or float32 at src_rate."""
        a = np.asarray(pcm)
        fmt = 1 if a.dtype == np.int16 else 0
        if fmt == 0:
            a = a.astype(np.float32, copy=False)
        a = np.ascontiguousarray(a)
        ch = 1 if a.ndim == 1 else a.shape[1]
        n_out = int(self._lib.bd_resample_out_len(a.shape[0], src_rate))
        _, _, P = frames_for(n_out, hop_frames)
        act = np.empty((P, self.n_classes), dtype=np.float32)
        if (n_out > getattr(self, "_reserved", 0) or a.nbytes > getattr(self, "_reserved_pcm", 0)) and not self._outstanding:
            self.reserve(n_out, a.nbytes, hop_frames)
        slot = self._acquire_slot(n_out)
        tk = self._register(slot, act, None, n_out, a)
        try:
            got = self.submit_pcm_ptr(slot, a.ctypes.data, fmt, ch, a.shape[0], src_rate, hop_frames, act.ctypes.data)
        except Exception:
            tk._done = True
            self._release(tk)
            raise
        assert got == P
        return tk

    def reserve(self, n_samples: int, pcm_bytes: int = 0, hop_frames: int = 96):
        """Pre-size every free slot for chunks of this size (no allocation while chunks are in flight)."""
        self._check(self._lib.bd_reserve_slots(self._h, int(n_samples), int(pcm_bytes), int(hop_frames)), "bd_reserve_slots")
        self._reserved = max(getattr(self, "_reserved", 0), int(n_samples))
        self._reserved_pcm = max(getattr(self, "_reserved_pcm", 0), int(pcm_bytes))

    def debug_stats(self) -> str:
        buf = C.create_string_buffer(512)
        self._check(self._lib.bd_debug_stats(self._h, buf, 512), "bd_debug_stats")
        return buf.value.decode()

    def trace(self, on: bool):
        """True: start the slot-API timeline.  False: stop; returns [(kind, a, b, host_ms, device_ms), ...]."""
        if on:
            self._check(self._lib.bd_trace(self._h, 1, None, 0), "bd_trace")
            return None
        buf = C.create_string_buffer(1 << 20)
        self._check(self._lib.bd_trace(self._h, 0, buf, len(buf)), "bd_trace")
        out = []
        for line in buf.value.decode().splitlines():
            k, a, b, h, d = line.split()
            out.append((int(k), int(a), int(b), float(h), float(d)))
        return out

    def set_auto_flush(self, on: bool):
        """False: submitted chunks are only launched by flush() / a ticket's wait() (deterministic batches)."""
        self._check(self._lib.bd_set_auto_flush(self._h, 1 if on else 0), "bd_set_auto_flush")

    def flush(self):
        """Launch every chunk that is still waiting for company (does not wait for results)."""
        self._check(self._lib.bd_flush(self._h), "bd_flush")

    def drain(self):
        """Complete every outstanding ticket (results stay readable through the tickets)."""
        while True:
            with self._tk_lock:
                tk = self._outstanding[0] if self._outstanding else None
            if tk is None:
                return
            tk.wait()

    @property
    def batch_stats(self) -> tuple[int, int]:
        """(CNN passes launched through the slot API, chunks they carried)."""
        b, c = C.c_int64(), C.c_int64()
        self._check(self._lib.bd_batch_stats(self._h, C.byref(b), C.byref(c)), "bd_batch_stats")
        return b.value, c.value

    def predict_ptr(self, host_ptr: int, n: int, hop_frames: int, act_ptr: int, emb_ptr: int | None = None) -> int:
        npat = C.c_int64()
        self._check(self._lib.bd_predict_host(self._h, host_ptr, n, hop_frames, act_ptr, emb_ptr, C.byref(npat)),
                    "bd_predict_host")
        return npat.value

    def submit_ptr(self, slot: int, host_ptr: int, n: int, hop_frames: int, act_ptr: int,
                   emb_ptr: int | None = None) -> int:
        npat = C.c_int64()
        self._check(self._lib.bd_submit_host(self._h, slot, host_ptr, n, hop_frames, act_ptr, emb_ptr, C.byref(npat)),
                    "bd_submit_host")
        return npat.value

    def submit_pcm_ptr(self, slot: int, pcm_ptr: int, fmt: int, channels: int, n_frames: int, src_rate: int,
                       hop_frames: int, act_ptr: int, emb_ptr: int | None = None) -> int:
        """Raw decoded PCM in (int16/float32, interleaved), activations out; downmix + resample run on the GPU."""
        npat = C.c_int64()
        self._check(self._lib.bd_submit_pcm_host(self._h, slot, pcm_ptr, fmt, channels, n_frames, src_rate, hop_frames,
                                                 act_ptr, emb_ptr, C.byref(npat)), "bd_submit_pcm_host")
        return npat.value

    def predict_pcm(self, pcm: np.ndarray, src_rate: int, hop_frames: int = 96) -> np.ndarray:
        """[n] or [n, channels] int16/float32 at src_rate -> [P, n_classes] (synchronous convenience wrapper)."""
        a = np.asarray(pcm)
        fmt = 1 if a.dtype == np.int16 else 0
        if fmt == 0:
            a = a.astype(np.float32, copy=False)
        a = np.ascontiguousarray(a)
        ch = 1 if a.ndim == 1 else a.shape[1]
        return self.submit_pcm(a, src_rate, hop_frames).result()

    def wait(self, slot: int):
        self._check(self._lib.bd_wait(self._h, slot), "bd_wait")

    def synchronize(self):
        self._check(self._lib.bd_synchronize(self._h), "bd_synchronize")

    def predict_device_ptr(self, d_samples: int, n: int, hop_frames: int, d_act: int, d_emb: int | None = None) -> int:
        npat = C.c_int64()
        self._check(self._lib.bd_predict_device(self._h, d_samples, n, hop_frames, d_act, d_emb, C.byref(npat)),
                    "bd_predict_device")
        return npat.value

    def bench_device_ptr(self, d_samples: int, n: int, hop_frames: int, d_act: int, steps: int) -> float:
        """milliseconds (CUDA events on the engine's compute stream) for `steps` passes over one resident chunk."""
        ms = C.c_float()
        self._check(self._lib.bd_bench_device(self._h, d_samples, n, hop_frames, d_act, steps, C.byref(ms)),
                    "bd_bench_device")
        return float(ms.value)

    def profile_device_ptr(self, d_samples: int, n: int, hop_frames: int = 96) -> dict:
        ms = (C.c_float * 29)()
        cnt = (C.c_int64 * 29)()
        self._check(self._lib.bd_profile_device(self._h, d_samples, n, hop_frames, ms, cnt), "bd_profile_device")
        groups = {"frontend": [0], "conv1": [1], "depthwise": list(range(2, 15)), "pointwise": list(range(15, 28)),
                  "pool_head": [28]}
        out = {s: {"ms": float(sum(ms[i] for i in idx)), "launches": int(sum(cnt[i] for i in idx))}
               for s, idx in groups.items()}
        out["layers"] = {f"L{i + 2}": {"dw_ms": float(ms[2 + i]), "pw_ms": float(ms[15 + i]),
                                       "dw_launches": int(cnt[2 + i]), "pw_launches": int(cnt[15 + i])}
                         for i in range(13)}
        return out

    # ------------------------------------------------------------------ resampler
    def resample(self, samples: np.ndarray, src_rate: int) -> np.ndarray:
        """[n] or [n, channels], float32 or int16 -> float32 mono at 16 kHz."""
        a = np.asarray(samples)
        if a.dtype == np.int16:
            fmt = 1
        else:
            a = a.astype(np.float32, copy=False)
            fmt = 0
        a = np.ascontiguousarray(a)
        ch = 1 if a.ndim == 1 else a.shape[1]
        n = a.shape[0]
        no = int(self._lib.bd_resample_out_len(n, src_rate))
        out = np.empty(no, dtype=np.float32)
        got = C.c_int64()
        self._check(self._lib.bd_resample_host(self._h, _ptr(a), fmt, ch, n, src_rate, _ptr(out), no, C.byref(got)),
                    "bd_resample_host")
        return out[:got.value]

    def resample_device_ptr(self, d_in: int, fmt: int, channels: int, n_frames: int, src_rate: int, d_out: int,
                            out_capacity: int) -> int:
        """Device-resident PCM in (fmt 0 float32 / 1 int16, interleaved), float32 mono at 16 kHz out; synchronous."""
        got = C.c_int64()
        self._check(self._lib.bd_resample_device(self._h, d_in, fmt, channels, n_frames, src_rate, d_out, out_capacity,
                                                 C.byref(got)), "bd_resample_device")
        return got.value

    # ------------------------------------------------------------------ test hooks
    def debug_logmel(self, samples: np.ndarray, n_frames: int) -> np.ndarray:
        x = np.ascontiguousarray(samples, dtype=np.float32)
        out = np.empty((n_frames, 64), dtype=np.float32)
        self._check(self._lib.bd_debug_logmel(self._h, _ptr(x), x.size, n_frames, _ptr(out)), "bd_debug_logmel")
        return out

    def debug_pw_gemm(self, A: np.ndarray, Wt: np.ndarray, bias: np.ndarray, precision: str, block_n: int = 0):
        A = np.ascontiguousarray(A, dtype=np.float32)
        Wt = np.ascontiguousarray(Wt, dtype=np.float32)
        bias = np.ascontiguousarray(bias, dtype=np.float32)
        M, K = A.shape
        N = Wt.shape[0]
        out = np.empty((M, N), dtype=np.float32)
        self._check(self._lib.bd_debug_pw_gemm(self._h, _ptr(A), _ptr(Wt), _ptr(bias), M, N, K, PRECISION[precision],
                                               block_n, _ptr(out)), "bd_debug_pw_gemm")
        return out

    def debug_stage(self, samples: np.ndarray, stage: int, hop_frames: int = 96, capacity: int = 1 << 26):
        x = np.ascontiguousarray(samples, dtype=np.float32)
        out = np.empty(capacity, dtype=np.float32)
        got = C.c_int64()
        self._check(self._lib.bd_debug_stage(self._h, _ptr(x), x.size, hop_frames, stage, _ptr(out), capacity,
                                             C.byref(got)), "bd_debug_stage")
        return out[:got.value].copy()
