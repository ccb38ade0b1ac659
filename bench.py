#!/usr/bin/env python
"""bench.py -- audio-hours processed per second through the buzzdetect inference hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path (frontend -> YAMNet MobileNet-v1 -> model_general_v3 head) over one synthetic
1-hour 16 kHz mono recording per GPU (BASELINE.json configs[1]; 57.6 M samples, larger than the 126 MB L2, so every
step re-reads its audio from HBM).  Files shard across GPUs with no collective (weak scaling).

  value      : audio-hours / s with the audio already resident in HBM, CUDA events on the engine's stream
  e2e        : the same metric through the reference-facing PLUGIN: load_model('model_general_v3') and one predict call
               per 199.68 s chunk (the reference's default chunking) from ONE inferer thread, results.numpy() on a
               writer thread (src/inference/worker.py:71-92, src/write/worker.py:67-70).  Chunks lie in pinned host
               memory as int16 PCM -- the file pipeline's feed -- and every chunk's host->device copy and every
               result's device->host copy are inside the timed region.  (= e2e_pcm16)
  e2e_plugin : the same with float32 samples through predict() (the reference streamer's dtype; PCIe-bound)
  e2e_plugin_pageable / e2e_slots : pageable numpy input; the raw bd_submit_host / bd_wait C-ABI leg
  configs    : BASELINE configs[4] (half hop) and configs[2] (44.1 kHz stereo int16 streamed in chunks) on the same path
  roofline / rooflines : per kernel family against MEASURED_PEAKS.json (algorithmic bytes or flops / CUDA-event time)
  cpu_baseline : the oracle (numpy/torch-CPU restatement of the reference's TensorFlow path) on this box's cores
  --impl reference : the CPU oracle on all host cores over the same workload (TensorFlow/librosa cannot be installed
               here and the YAMNet blob is missing from the checkout)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the YAMNet variables blob is absent from the reference checkout (.MISSING_LARGE_BLOBS): the bench runs the seeded
# synthetic network of the exact architecture unless BUZZ_YAMNET_WEIGHTS points at the real blob ("weights" in config)
os.environ.setdefault("BUZZ_B200_ALLOW_SYNTHETIC", "1")

SR = 16000
HOUR_SAMPLES = 3600 * SR
PATCHES_PER_HOUR = 3750
HOP_FRAMES = 96
# algorithmic work per 0.96 s patch (BASELINE.md section 3 / SURVEY.md section 8d)
PW_FLOP_PER_PATCH = 132_120_576
DW_BYTES_PER_PATCH = 2_445_312
FRONTEND_BYTES_PER_PATCH = 86_016
FE_WARP_INSTR_PER_FRAME = 417          # logmel2_kernel: smsp__inst_executed.sum / frames (ncu, round 2: 150.3 M / 360,000)
TOTAL_FLOP_PER_PATCH = 137_289_728
METRIC = "audio-hours processed/sec (realtime factor) at 1/2/4/8 B200 vs host-CPU ref"
UNIT = "audio-hours/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions (B200_PROFILING.md), in-process through NVML.

    A polling `nvidia-smi -lms 50` process is what the recipe shows, and it is fine around one long kernel, but its
    queries take driver locks: with ~100 small chunks and ~30 kernel launches per pass in flight it stretched the
    plugin end-to-end legs six-fold (measured: 17 vs 217 audio-hours/s).  pynvml asks for three values every 20 ms
    from a thread of this process instead; nvidia-smi stays as the fallback when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.samples = []          # (sm_mhz, reasons bitmask, power_w, inside a timed region or its warm-up)
        self.active = False
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self.how = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._nv = (pynvml, h)
            self.how = "pynvml, 20 ms"

            def loop():
                nv, hh = self._nv
                while not self._stop.is_set():
                    try:
                        mhz = float(nv.nvmlDeviceGetClockInfo(hh, nv.NVML_CLOCK_SM))
                        rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(hh)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                            else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(hh))
                        pw = nv.nvmlDeviceGetPowerUsage(hh) / 1000.0
                        self.samples.append((mhz, rs, pw, self.active))
                    except Exception:                                  # noqa: BLE001 -- diagnostics only
                        pass
                    self._stop.wait(0.02)

            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
            return
        except Exception:                                              # noqa: BLE001 -- fall back to nvidia-smi
            self._nv = None
        try:
            self.how = "nvidia-smi -lms 200"
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            nv = self._nv[0]
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            # samples taken while a leg (warm-up + timed region) was running; the GPU idles between the legs while
            # host buffers are prepared
            busy = [x for x in self.samples if x[3]] or self.samples
            reasons = sorted(nm for nm, b in bits.items() if any(x[1] & b for x in busy))
            sm = [x[0] for x in busy]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "samples_total": len(self.samples), "power_w_max": max((x[2] for x in self.samples), default=None),
                    "reasons": reasons, "how": self.how}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "how": self.how}


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ----------------------------------------------------------------------------------------------- CPU oracle timing
def time_oracle(steps: int, warmup: int, hours_per_step: float = 1.0, chunk_s: float = 199.68, seed: int = 1000,
                budget_s: float = 200.0):
    """The restated reference path on the host cores: one synthetic recording of `hours_per_step` per step, cut into
    the reference's default 199.68 s chunks (src/analyze.py:102-111), every chunk through predict(), all threads."""
    import numpy as np
    import torch
    from buzzdetect_b200 import weights as W
    from oracle import yamnet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    variables, prov = W.resolve_yamnet(verify=False)
    hk, hb = W.load_head()
    mel = W.load_mel()
    n = int(round(hours_per_step * HOUR_SAMPLES))
    base = O.synth_audio(60 * SR, seed=seed)
    x = np.tile(base, -(-n // base.size))[:n]
    chunk_n = int(round(chunk_s * SR))
    chunks = [x[o:o + chunk_n] for o in range(0, n, chunk_n)]

    # bounded: one chunk is timed first; if the whole run would exceed the budget on this host, a step covers only the
    # first k chunks of the recording (said in `sample`; rates are per audio actually processed)
    t0 = time.perf_counter()
    O.predict(chunks[0], variables, mel, hk, hb, HOP_FRAMES)
    t_chunk = time.perf_counter() - t0
    k = len(chunks)
    if (steps + warmup) * k * t_chunk > budget_s:
        k = max(1, int(budget_s / ((steps + warmup) * t_chunk)))
    used = chunks[:k]
    audio_h = sum(c.size for c in used) / HOUR_SAMPLES

    def one_step():
        for c in used:
            O.predict(c, variables, mel, hk, hb, HOP_FRAMES)

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = time.perf_counter() - t0
    hours = steps * audio_h
    what = (f"the whole {hours_per_step:g}-hour recording ({len(chunks)} chunks of {chunk_s} s)" if k == len(chunks) else
            f"the first {k} of {len(chunks)} chunks of {chunk_s} s of the {hours_per_step:g}-hour recording (time budget)")
    return hours / dt, dt, cores, (f"{steps} step(s) over {what}, reference default chunking, {warmup} warm-up step(s); "
                                  f"oracle = torch-CPU/numpy restatement on {cores} threads, weights {prov.split(':')[0]}")


def run_reference(args, rank, world):
    if rank != 0:
        return
    from buzzdetect_b200 import probe
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    value, dt, cores, sample = time_oracle(steps, warmup, args.hours)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.hours),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_runtime": probe.reference_runtime(),
        "note": "TensorFlow/librosa are absent from this image and the YAMNet blob is absent from the reference "
                "checkout, so the reference arm is the CPU oracle port on all host cores (DESIGN.md section 8)",
    }
    print(json.dumps(line), flush=True)


def workload_config(hours):
    return {"workload": f"synthetic {hours:g}-hour 16 kHz mono recording per GPU, YAMNet + model_general_v3, hop "
                        f"{HOP_FRAMES / 96:g} (BASELINE configs[{1 if HOP_FRAMES == 96 else 4}])",
            "chunk_s": 199.68}


def bind_to_gpu_numa_node(dev: int):
    """Best effort: run this rank (and allocate its pinned buffers) on the CPUs of the NUMA node the GPU hangs off, so
    that eight ranks do not pull their host->device traffic through one socket.  Returns a note for the JSON line."""
    try:
        out = subprocess.run(["nvidia-smi", "-i", str(dev), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip()
        if not out:
            return "numa: no pci bus id"
        bus = out.lower()                                         # 00000000:1b:00.0 -> 0000:1b:00.0
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa: single node"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return f"numa: node {node} has no allowed cpus"
        os.sched_setaffinity(0, allowed)
        return f"numa: bound to node {node} ({len(allowed)} cpus)"
    except Exception as exc:                                          # noqa: BLE001 -- diagnostics only
        return f"numa: not bound ({type(exc).__name__})"


# ----------------------------------------------------------------------------------------------- configs 3 and 5
def run_extra_configs(args, rank, world, dev, eng, model, d_x, d_act, n, hv, use_dist, dist, capi, np, torch):
    """BASELINE configs[4] (yamnet_k2 at hop 0.5) and configs[2] (44.1 kHz stereo int16 streamed in 199.68 s chunks)
    through the same plugin path, one recording-hour each; reported beside the headline, never instead of it."""
    import queue
    from buzzdetect_b200.inference.models import load_model
    from oracle import yamnet_oracle as O
    out = {}
    hours = n / HOUR_SAMPLES

    def through_plugin(m, feed_chunks, rate, steps):
        def run(k):
            q = queue.Queue(maxsize=args.slots)

            def writer():
                while True:
                    item = q.get()
                    if item is None:
                        return
                    item.numpy()

            th = threading.Thread(target=writer)
            th.start()
            for _ in range(k):
                for c in feed_chunks:
                    q.put(m.predict_pcm(c, rate))
            q.put(None)
            th.join()

        run(2)
        m.model.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run(steps)
        m.model.synchronize()
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if use_dist:
            dist.barrier()
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * hours * steps / float(tt.item())

    steps = max(1, min(args.steps, 3))
    # ---- config 5 / configs[4]: half hop (7499 patches per audio-hour): device-resident and through the plugin
    _, _, P48 = capi.frames_for(n, 48)
    d_act48 = torch.empty((P48, eng.n_classes), dtype=torch.float32, device="cuda")
    for _ in range(2):
        eng.predict_device_ptr(d_x.data_ptr(), n, 48, d_act48.data_ptr())
    ms48 = eng.bench_device_ptr(d_x.data_ptr(), n, 48, d_act48.data_ptr(), steps)
    t48 = torch.tensor([ms48], dtype=torch.float64, device="cuda")
    if use_dist:
        dist.all_reduce(t48, op=dist.ReduceOp.MAX)
    m48 = load_model("model_general_v3", framehop_prop=0.5, initialize=True)
    chunk_n = int(round(args.chunk_s * SR))
    pcm16 = capi.pinned_empty(n, np.int16)
    pcm16[:] = np.clip(np.rint(hv * 32768.0), -32768, 32767).astype(np.int16)
    feed = [pcm16[o:o + chunk_n] for o in range(0, n, chunk_n)]
    out["hop48"] = {"value": world * hours * steps / (float(t48.item()) / 1e3), "e2e": through_plugin(m48, feed, SR, steps),
                    "unit": UNIT, "patches_per_step_per_gpu": P48,
                    "what": "BASELINE configs[4]: yamnet_k2 halfhop (framehop_prop 0.5) + model_general_v3; e2e = plugin "
                            "predict_pcm per 199.68 s chunk, pinned int16"}
    m48.model.close()
    del d_act48, feed, pcm16
    # ---- config 3 / configs[2]: 44.1 kHz stereo int16 streamed in the reference's 199.68 s chunks (one hour of the
    # day-long recording per step): raw PCM over PCIe, downmix + resample + path on the device
    sr3 = 44100
    n3 = int(round(hours * 3600 * sr3))
    base = O.synth_audio(60 * sr3, seed=2000 + rank, sr=sr3)
    b16 = np.clip(np.rint(base * 32768.0), -32768, 32767).astype(np.int16)
    pcm3 = capi.pinned_empty((n3, 2), np.int16)
    for off in range(0, n3, b16.size):
        k = min(b16.size, n3 - off)
        pcm3[off:off + k, 0] = b16[:k]
        pcm3[off:off + k, 1] = np.roll(b16, 37)[:k]
    chunk3 = int(args.chunk_s * sr3)                       # int(chunk[1] * sr): src/stream/worker.py:110-112
    feed3 = [pcm3[o:o + chunk3] for o in range(0, n3, chunk3)]
    out["cfg3_44k1_stereo_int16"] = {
        "e2e": through_plugin(model, feed3, sr3, steps), "unit": UNIT, "h2d_bytes_per_step": int(pcm3.nbytes),
        "chunks_per_step": len(feed3),
        "what": "BASELINE configs[2], one hour of the recording per step: 44.1 kHz stereo int16 in pinned host memory, "
                "199.68 s chunks through predict_pcm (downmix + tcgen05 resampler + path on the device); the resume run "
                "is a parity test (tests/test_writer_pipeline.py), not a timing"}
    del feed3, pcm3
    # ---- config 1 / configs[0]: the reference's own audio_in/testbuzz.mp3 (6.55 s, 32 kHz mono; PCM fixture of its decode):
    # one file = one 7-patch chunk, so this is the latency of a single small pass, resampler included
    fx = os.path.join(ROOT, "tests", "golden", "testbuzz_32k_s16.wav")
    if rank == 0 and os.path.exists(fx):
        import wave
        with wave.open(fx, "rb") as w:
            sr1, nfr = w.getframerate(), w.getnframes()
            raw = np.frombuffer(w.readframes(nfr), dtype=np.int16).copy()
        pin1 = capi.pinned_empty(raw.size, np.int16)
        pin1[:] = raw
        for _ in range(5):
            rows = model.predict_pcm(pin1, sr1).numpy().shape[0]
        t0 = time.perf_counter()
        reps = 50
        for _ in range(reps):
            model.predict_pcm(pin1, sr1).numpy()
        dt = (time.perf_counter() - t0) / reps
        out["cfg1_testbuzz"] = {"ms_per_file": dt * 1e3, "value": (nfr / sr1 / 3600.0) / dt, "unit": UNIT, "rows": int(rows),
                                "what": "BASELINE configs[0]: testbuzz.mp3 as decoded PCM (32 kHz mono int16, 6.55 s), one "
                                        "synchronous predict_pcm(...).numpy() per file: latency of one 7-patch pass"}
    return out


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local):
    import numpy as np
    import torch
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        # NCCL announces its version on stdout when NCCL_DEBUG asks for it; stdout carries exactly one JSON line, so
        # anything the communicator setup prints goes to stderr instead
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            t0_ = torch.zeros(1, device="cuda")
            dist.all_reduce(t0_)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    from buzzdetect_b200 import capi
    from oracle import yamnet_oracle as O            # synthetic audio generator + cpu_baseline only

    dev = local if use_dist else 0
    torch.cuda.set_device(dev)
    numa_note = bind_to_gpu_numa_node(dev) if use_dist else "numa: single rank, not bound"
    eng = capi.Engine(device=dev, precision=args.precision, early_patches=args.early, late_patches=args.late,
                      n_slots=4, fuse_mask=args.fuse_mask)
    hours_per_step = args.hours
    n = int(round(hours_per_step * HOUR_SAMPLES))
    # one hour of synthetic audio: 60 s of the oracle's generator tiled with per-rank seed (host I/O is not measured)
    base = O.synth_audio(60 * SR, seed=1000 + rank)
    host = torch.empty(n, dtype=torch.float32).pin_memory()
    hv = host.numpy()
    for off in range(0, n, base.size):
        m = min(base.size, n - off)
        hv[off:off + m] = base[:m]
    _, _, P = capi.frames_for(n, HOP_FRAMES)
    d_x = host.cuda(non_blocking=False)
    d_act = torch.empty((P, eng.n_classes), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()

    # ---------------- device-resident throughput ("value")
    # nvidia-smi needs ~100 ms to start reporting, longer than one timed region, so the sampler runs from the warm-up
    # of the device-resident leg to the end of the timed end-to-end leg: every sample is taken under load.
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    sampler.active = True
    for _ in range(max(args.warmup, 3)):
        eng.predict_device_ptr(d_x.data_ptr(), n, HOP_FRAMES, d_act.data_ptr())
    if use_dist:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = eng.launch_count
    ms = eng.bench_device_ptr(d_x.data_ptr(), n, HOP_FRAMES, d_act.data_ptr(), args.steps)
    torch.cuda.synchronize()
    launches = eng.launch_count - l0
    sampler.active = False
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if use_dist:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * hours_per_step * args.steps / (ms_max / 1000.0)

    # ---------------- end to end through the reference-facing plugin (load_model -> predict per chunk)
    # Thread layout of the reference: ONE inferer thread calls model.predict(chunk) per 199.68 s chunk
    # (src/inference/worker.py:71-92) and hands the result to the writer thread, which calls results.numpy()
    # (src/write/worker.py:67-70).  The chunks lie in PINNED host memory, as the pinned-buffer streamer leaves them.
    # Host->device copy of every chunk and device->host copy of every result are inside the timed region.
    import queue
    os.environ["BUZZ_B200_DEVICE"] = str(dev)
    os.environ["BUZZ_B200_PRECISION"] = args.precision
    os.environ["BUZZ_B200_SLOTS"] = str(args.slots)
    from buzzdetect_b200.inference.models import load_model
    framehop_prop = HOP_FRAMES / 96.0
    model = load_model("model_general_v3", framehop_prop=framehop_prop, initialize=True)
    pcm16 = capi.pinned_empty(n, np.int16, write_combined=args.wc)
    pcm16[:] = np.clip(np.rint(hv * 32768.0), -32768, 32767).astype(np.int16)
    chunk_n = int(round(args.chunk_s * SR))
    bounds = [(o, min(chunk_n, n - o)) for o in range(0, n, chunk_n)]

    def plugin_run(steps, feed):
        """`steps` recordings back to back through the plugin; returns patches written by the writer thread."""
        q = queue.Queue(maxsize=args.slots)
        done = {"rows": 0}

        def writer():
            while True:
                item = q.get()
                if item is None:
                    return
                done["rows"] += item.numpy().shape[0]

        th = threading.Thread(target=writer)
        th.start()
        t_sub = time.perf_counter()
        for _ in range(steps):
            for (o, m) in bounds:
                if feed == "f32":
                    q.put(model.predict(hv[o:o + m]))
                elif feed == "pcm16":
                    q.put(model.predict_pcm(pcm16[o:o + m], SR))
                else:
                    q.put(model.predict(pageable[o:o + m]))
        plugin_run.submit_s = time.perf_counter() - t_sub
        q.put(None)
        th.join()
        return done["rows"]

    def timed_plugin(feed, steps):
        sampler.active = True
        plugin_run(max(args.warmup, 3), feed)
        model.model.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()
        b0, c0 = model.model.batch_stats
        t0 = time.perf_counter()
        rows = plugin_run(steps, feed)
        model.model.synchronize()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        sampler.active = False
        b1, c1 = model.model.batch_stats
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if use_dist:
            dist.barrier()
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        assert rows == steps * sum(capi.frames_for(m, HOP_FRAMES)[2] for _, m in bounds)
        return {"value": world * hours_per_step * steps / float(tt.item()), "unit": UNIT,
                "chunks_per_pass": round((c1 - c0) / max(b1 - b0, 1), 2), "passes": int(b1 - b0),
                "wall_s": dt, "inferer_loop_s": plugin_run.submit_s}

    def with_h2d_rate(d):
        """host->device GB/s this leg sustained, summed over the ranks (what a shared host has to deliver)."""
        if d.get("h2d_bytes_per_step"):
            d["h2d_gbs_all_gpus"] = d["h2d_bytes_per_step"] * d["value"] / hours_per_step / 1e9
        return d

    d2h_bytes = sum(capi.frames_for(m, HOP_FRAMES)[2] for _, m in bounds) * eng.n_classes * 4
    e2e_pcm16 = timed_plugin("pcm16", args.steps)
    e2e_pcm16.update({"h2d_bytes_per_step": n * 2, "d2h_bytes_per_step": d2h_bytes, "chunk_s": args.chunk_s,
                      "chunks_per_step": len(bounds), "entry": "load_model('model_general_v3').predict_pcm(int16 chunk, 16000) "
                      "per chunk from one inferer thread, results.numpy() on a writer thread; pinned int16 PCM (what a WAV "
                      "streamer holds), converted on the device"})
    e2e_plugin = timed_plugin("f32", args.steps)
    e2e_plugin.update({"h2d_bytes_per_step": n * 4, "d2h_bytes_per_step": d2h_bytes, "chunk_s": args.chunk_s,
                       "chunks_per_step": len(bounds), "entry": "load_model('model_general_v3').predict(float32 chunk) per chunk, "
                       "pinned float32 samples (the reference streamer's dtype): 230 MB per audio-hour over PCIe"})
    with_h2d_rate(e2e_pcm16)
    with_h2d_rate(e2e_plugin)
    pageable = np.array(hv, copy=True)
    e2e_pageable = timed_plugin("pageable", max(1, args.steps // 2))
    e2e_pageable.update({"h2d_bytes_per_step": n * 4, "entry": "the same with pageable numpy input (host->device copies are "
                         "staged synchronously by the driver)"})
    del pageable
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- the raw C-ABI slot leg (bd_submit_host / bd_wait, no plugin objects), longer chunks
    slot_chunk_n = int(round(args.slot_chunk_s * SR))
    chunks = [(o, min(slot_chunk_n, n - o)) for o in range(0, n, slot_chunk_n)]
    outs = [torch.empty((capi.frames_for(m, HOP_FRAMES)[2], eng.n_classes), dtype=torch.float32).pin_memory()
            for (_, m) in chunks]
    n_slots = 16
    slot_eng = capi.Engine(device=dev, precision=args.precision, early_patches=args.early, late_patches=args.late,
                           n_slots=n_slots, fuse_mask=args.fuse_mask)

    def slots_run(steps):
        j = 0
        for _ in range(steps):
            for i, (o, m) in enumerate(chunks):
                sl = j % n_slots
                j += 1
                slot_eng.wait(sl)
                slot_eng.submit_ptr(sl, host.data_ptr() + 4 * o, m, HOP_FRAMES, outs[i].data_ptr())
        for sl in range(n_slots):
            slot_eng.wait(sl)

    slots_run(max(args.warmup, 3))
    slot_eng.synchronize()
    if use_dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    slots_run(args.steps)
    slot_eng.synchronize()
    torch.cuda.synchronize()
    tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if use_dist:
        dist.barrier()
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_slots = {"value": world * hours_per_step * args.steps / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": n * 4,
                 "chunk_s": args.slot_chunk_s, "slots": n_slots, "entry": "bd_submit_host / bd_wait, pinned float32"}
    slot_eng.close()

    # ---------------- BASELINE configs 3 and 5 through the same plugin path
    extra_configs = {}
    if not args.no_configs:
        extra_configs = run_extra_configs(args, rank, world, dev, eng, model, d_x, d_act, n, hv, use_dist,
                                          dist if use_dist else None, capi, np, torch)

    # ---------------- the stage in front of the path for non-16 kHz input (BASELINE configs[2]): downmix + resample of
    # one audio-hour of 44.1 kHz stereo int16 PCM already in HBM (synchronous C-ABI call, wall clock around it)
    resample_stage = None
    if not args.no_resample:
        n441 = 3600 * 44100
        g_ = torch.Generator(device="cuda").manual_seed(7 + rank)
        pcm441 = (torch.randn((n441, 2), device="cuda", generator=g_) * 3000).to(torch.int16)
        n_rs = capi.load_library().bd_resample_out_len(n441, 44100)
        d_rs = torch.empty(int(n_rs) + 256, dtype=torch.float32, device="cuda")
        for _ in range(2):
            eng.resample_device_ptr(pcm441.data_ptr(), 1, 2, n441, 44100, d_rs.data_ptr(), d_rs.numel())
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            eng.resample_device_ptr(pcm441.data_ptr(), 1, 2, n441, 44100, d_rs.data_ptr(), d_rs.numel())
        rs_ms = (time.perf_counter() - t0) / reps * 1e3
        rs_bytes = n441 * 2 * 2 + int(n_rs) * 4
        resample_stage = {"ms_per_audio_hour": rs_ms, "input": "44.1 kHz stereo int16, device-resident",
                          "algorithmic_gbs": rs_bytes / (rs_ms / 1e3) / 1e9, "frac_hbm": rs_bytes / (rs_ms / 1e3) / 1e9 / measured_peaks()["hbm_gbs"],
                          "kernel": "resample_tc_kernel (tcgen05 GEMM over blocks of 160 outputs) + resample_kernel tail"}
        del pcm441, d_rs

    # ---------------- per-kernel-family device times (one extra un-graphed pass, CUDA events around every launch)
    prof = eng.profile_device_ptr(d_x.data_ptr(), n, HOP_FRAMES)
    peaks = measured_peaks()
    from buzzdetect_b200.weights import LAYERS
    mma_factor = {"fp16x3": 3, "fp16f8": 2, "fp16": 1, "fp32": 0}[args.precision]
    per_layer = {}
    def _f():
        return {"ms": 0.0, "launches": 0, "flop": 0.0, "bytes": 0.0}
    # kernel families; a fused separable layer is HBM-bound while K <= 256 (layers 3-6) and tensor-bound at K = 512
    fam = {"pw_gemm_kernel": _f(), "sep_fused3_kernel[layers 3-7, hbm]": _f(),
           "sep_fused3_kernel[layers 8-12, tensor]": _f(), "depthwise_kernel": _f()}
    for i, (kind, stride, cin, cout, H, W) in enumerate(LAYERS[1:]):
        v = prof["layers"][f"L{i + 2}"]
        ho, wo = H // stride, W // stride
        dw_bytes = (H * W * cin + ho * wo * cin) * 4.0 * P              # fp32 in + (hi+lo fp16 = 4 B) out
        pw_flop = 2.0 * ho * wo * cin * cout * P
        pw_bytes = (ho * wo * cin + ho * wo * cout) * 4.0 * P
        fused = v["dw_launches"] == 0 and v["pw_launches"] > 0 and i + 2 != 2
        per_layer[f"L{i + 2}"] = {
            "dw_ms": round(v["dw_ms"], 4), "dw_gbs": round(dw_bytes / max(v["dw_ms"], 1e-9) / 1e6, 1) if v["dw_ms"] else None,
            "pw_ms": round(v["pw_ms"], 4), "pw_tflops": round(pw_flop / max(v["pw_ms"], 1e-9) / 1e9, 1) if v["pw_ms"] else None,
            "fused": fused, "K": cin, "N": cout, "M": ho * wo * P}
        if v["dw_launches"]:
            f = fam["depthwise_kernel"]
            f["ms"] += v["dw_ms"]; f["launches"] += v["dw_launches"]; f["bytes"] += dw_bytes
        if v["pw_launches"]:
            if fused:
                nm = "sep_fused3_kernel[layers 8-12, tensor]" if cin >= 512 else "sep_fused3_kernel[layers 3-7, hbm]"
            else:
                nm = "pw_gemm_kernel"
            f = fam[nm]
            f["ms"] += v["pw_ms"]; f["launches"] += v["pw_launches"]; f["flop"] += pw_flop
            f["bytes"] += ((H * W * cin + ho * wo * cout) * 4.0 * P) if fused else pw_bytes
    l12 = prof["layers"]["L2"]["pw_launches"] == 0 and prof["layers"]["L2"]["dw_launches"] == 0
    if l12:     # layers 1+2 in one kernel: log-mel patch in, layer-2 output out
        fam["l12_fused2_kernel"] = {"ms": prof["conv1"]["ms"], "launches": prof["conv1"]["launches"], "flop": 0.0,
                                    "bytes": (96 * 64 + 48 * 32 * 64) * 4.0 * P}
    else:       # layer 1 + layer-2 depthwise: log-mel patch in, hi/lo planes (4 B per element) out
        fam["conv1_dw2_kernel"] = {"ms": prof["conv1"]["ms"], "launches": prof["conv1"]["launches"], "flop": 0.0,
                                   "bytes": (96 * 64 + 48 * 32 * 32) * 4.0 * P}
    fam["logmel_kernel"] = {"ms": prof["frontend"]["ms"], "launches": prof["frontend"]["launches"], "flop": 0.0,
                            "bytes": FRONTEND_BYTES_PER_PATCH * float(P)}
    total_prof_ms = sum(prof[k]["ms"] for k in ("frontend", "conv1", "depthwise", "pointwise", "pool_head"))
    traffic_tab = {}
    import glob
    tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "traffic_*.json")))      # tags sort by round: r1 < r1e < r2
    if tfiles:
        with open(tfiles[-1]) as f:
            traffic_tab = json.load(f)
        traffic_tab["_file"] = os.path.basename(tfiles[-1])
    rooflines = {}
    for nm, f in fam.items():
        if f["ms"] <= 0:
            continue
        tens = nm.startswith("pw_gemm") or "tensor" in nm
        ach = (f["flop"] / (f["ms"] / 1e3) / 1e12) if tens else (f["bytes"] / (f["ms"] / 1e3) / 1e9)
        peak = peaks["bf16_tflops_sustained"] if tens else peaks["hbm_gbs"]
        rooflines[nm] = {"bound": "tensor" if tens else "hbm", "achieved": ach, "peak": peak,
                         "unit": "TFLOP/s" if tens else "GB/s", "frac": ach / peak, "ms": f["ms"],
                         "launches": f["launches"], "share_of_step": f["ms"] / total_prof_ms,
                         "algorithmic_per_launch": (f["flop"] if tens else f["bytes"]) / max(f["launches"], 1),
                         "avg_launch_ms": f["ms"] / max(f["launches"], 1),
                         "traffic": traffic_tab.get(nm, {}).get("dram_bytes_per_launch")}
        rooflines[nm]["traffic_source"] = (f"{traffic_tab.get('_file')}: {traffic_tab.get('_source')}"
                                           if rooflines[nm]["traffic"] else None)
        if tens:
            rooflines[nm]["executed_mma_factor"] = mma_factor
            rooflines[nm]["hbm_gbs"] = f["bytes"] / (f["ms"] / 1e3) / 1e9
    # what ncu saw of the SM's shared-memory data pipe (one 128-byte wavefront per clock, shared by LDS/STS/LDG hits and
    # the tensor core's operand fetch): the bound of the fused kernels that neither HBM nor the tensor pipe explains
    kfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_r*_kernels.csv")))
    if kfiles:
        import csv
        import re as _re
        with open(kfiles[-1]) as f:
            rows = list(csv.reader(f))
        hdr = rows[0]
        c_l = "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"
        c_t = "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"
        if c_l in hdr and c_t in hdr:
            acc = {}
            for r in rows[2:]:
                nm = r[1]
                if nm.startswith("sep_fused3_kernel"):
                    m = _re.search(r"<\s*\d+,\s*\d+,\s*(\d+)", nm)
                    key = "sep_fused3_kernel[layers 8-12, tensor]" if m and m.group(1) == "512" else "sep_fused3_kernel[layers 3-7, hbm]"
                elif nm.startswith("logmel"):
                    key = "logmel_kernel"
                else:
                    key = _re.sub(r"<.*", "", nm)
                try:
                    a = acc.setdefault(key, [0.0, 0.0, 0])
                    a[0] += float(r[hdr.index(c_l)]); a[1] += float(r[hdr.index(c_t)]); a[2] += 1
                except ValueError:
                    pass
            for key, (l, t, cnt) in acc.items():
                if key in rooflines and cnt:
                    rooflines[key]["smem_pipe_pct"] = {"lsu_wavefronts": l / cnt, "tensor_operand_wavefronts": t / cnt,
                                                      "note": "ncu, % of the one-wavefront-per-clock L1/shared data pipe; TMA writes "
                                                              "into shared memory use the same pipe and are in neither counter",
                                                      "source": os.path.basename(kfiles[-1])}
    dominant = max(rooflines, key=lambda k: rooflines[k]["ms"])
    roofline = dict(rooflines[dominant])
    roofline["kernel"] = dominant
    roofline["peak_source"] = peaks["source"] + (" bf16 sustained (kernel timed inside a long step)"
                                                if roofline["bound"] == "tensor" else " copy bandwidth")
    roofline["note"] = ("algorithmic flops (1x); the fp16x3 split executes 3 MMAs per algorithmic MMA, so the ceiling "
                        "of this mode is one third of the bf16 peak") if roofline["bound"] == "tensor" \
        else "algorithmic bytes (fp32 in + out)"
    stages = {k: prof[k] for k in ("frontend", "conv1", "depthwise", "pointwise", "pool_head")}

    # the frontend sits at the FP32-issue side of the ridge: report that bound beside HBM
    fe = rooflines.get("logmel_kernel")
    if fe is not None:
        frames = (P - 1) * HOP_FRAMES + 96
        warp_instr = frames * FE_WARP_INSTR_PER_FRAME
        sm_clock_hz = (clocks or {}).get("sm_mhz") or 1965.0
        ideal_ms = warp_instr / (148 * 4 * sm_clock_hz * 1e6) * 1e3
        fe["issue_bound"] = {"warp_instructions_per_frame": FE_WARP_INSTR_PER_FRAME, "ideal_ms_at_4_ipc_per_sm": ideal_ms,
                             "frac": ideal_ms / fe["ms"], "note": "512-point fp32 FFT + split + mel: executed warp instructions "
                             "per frame (ncu) against one instruction per scheduler per clock; the FMA pipe alone (about 300 of "
                             "them, FFMA = 1.04 clk per scheduler, tools/ubench/ffma2.cu) puts the floor near 0.14 ms per audio-hour"}

    if rank == 0:
        from buzzdetect_b200 import probe
        cpu_v, cpu_dt, cores, sample = time_oracle(steps=2, warmup=1, hours_per_step=hours_per_step, budget_s=30.0) \
            if (world == 1 and not args.no_cpu) else (None,) * 4
        cfgd = workload_config(hours_per_step)
        cfgd.update({"patches_per_step_per_gpu": P, "pointwise_precision": args.precision,
                     "weights": eng.weights_provenance.split(":")[0],
                     "l2": "input (230 MB/step) larger than L2; no explicit flush",
                     "early_patches": args.early, "late_patches": args.late, "fuse_mask": args.fuse_mask,
                     "plugin_slots": args.slots, "sharding": "one file per GPU, no collective", "host": numa_note})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.precision != "fp16" else "f16",
            "data": "synthetic",
            "config": cfgd,
            "realtime_factor": value * 3600.0,
            # headline: the plugin path fed the way the file pipeline feeds it (pinned int16 PCM chunks of 199.68 s)
            "e2e": dict(e2e_pcm16),
            "e2e_pcm16": e2e_pcm16,
            "e2e_plugin": e2e_plugin,
            "e2e_plugin_pageable": e2e_pageable,
            "e2e_slots": e2e_slots,
            "e2e_over_value": {"pcm16": e2e_pcm16["value"] / value, "plugin_f32": e2e_plugin["value"] / value},
            "configs": extra_configs,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "rooflines": rooflines,
            "stages": stages,
            "resample_stage": resample_stage,
            "layers": per_layer,
            "whole_path_tflops": TOTAL_FLOP_PER_PATCH * P * world * args.steps / (ms_max / 1000.0) / 1e12,
            "reference_runtime": probe.reference_runtime(),
        }
        if cpu_v is not None:
            line["cpu_baseline"] = {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    model.model.close()
    eng.close()


def main():
    global HOP_FRAMES
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("BUZZ_B200_PRECISION", "fp16x3"),
                    choices=["fp16x3", "fp16f8", "fp16", "fp32"])
    ap.add_argument("--hours", type=float, default=1.0, help="audio hours per step per GPU")
    ap.add_argument("--chunk-s", dest="chunk_s", type=float, default=199.68,
                    help="chunk length of the plugin end-to-end legs (reference default: 199.68 s)")
    ap.add_argument("--slot-chunk-s", dest="slot_chunk_s", type=float, default=599.04,
                    help="chunk length of the raw bd_submit_host leg")
    ap.add_argument("--slots", type=int, default=48, help="chunks in flight through the plugin (engine slots)")
    ap.add_argument("--no-configs", dest="no_configs", action="store_true", help="skip the configs 3 / 5 legs")
    ap.add_argument("--wc", action="store_true", help="experiment: write-combined pinned memory for the int16 feed")
    ap.add_argument("--early", type=int, default=0)
    ap.add_argument("--late", type=int, default=0)
    ap.add_argument("--fuse-mask", dest="fuse_mask", type=int, default=-1,
                    help="bit (L-2): run separable layer L as one fused depthwise+pointwise kernel (-1 = default)")
    ap.add_argument("--hop-frames", dest="hop_frames", type=int, default=96,
                    help="96 = framehop_prop 1 (headline), 48 = 0.5 (BASELINE configs[4], yamnet_k2 halfhop)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-resample", dest="no_resample", action="store_true", help="skip the 44.1 kHz resample-stage timing")
    args = ap.parse_args()
    HOP_FRAMES = args.hop_frames
    rank, world, local = dist_setup(args.gpus)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, world, local)


if __name__ == "__main__":
    main()
