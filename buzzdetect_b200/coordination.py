"""Coordinator, manifest and logger around PER-GPU analyze queues.

The reference runs streamer threads -> one bounded q_analyze -> analyzer threads -> q_write -> one writer, and owns
early/normal exit in a Coordinator that poisons every queue with one 'exit' sentinel per consumer
(src/pipeline/coordination.py:26-196).  This module keeps that protocol -- same getter/putter names, same sentinel,
same "first exit reason wins", same fully-analyzed bookkeeping -- and generalises the single q_analyze to one bounded
queue PER GPU, each drained by that GPU's inferer thread(s) (one engine per GPU, no collective: section 8e).  A chunk
goes to the GPU whose queue is shortest, so a slow device never stalls the others.

Also here, because they guard the same output folder:
  * the manifest lock (src/pipeline/manifest.py:13-85): settings that fix the schema of the result files must match
    what is already in the folder;
  * the log worker (src/pipeline/logger.py:23-65): one thread drains q_log into a file + console logger, PROGRESS level
    between DEBUG and INFO.
Workers (`StreamWorker`, `InferWorker`, `WriteWorker`) are the thread bodies; `run_analysis` wires them for a list of
files and a list of GPUs in ONE process (threads; the C ABI releases the GIL).  The CLI / GUI are not rebuilt.
"""
from __future__ import annotations

import json
import logging
import os
import threading
import time
from dataclasses import dataclass, field
from queue import Full, Queue

import numpy as np

from . import config as cfg, stream, write

EXIT = "exit"                     # the sentinel a worker receives instead of work (reference: coordination.py:10)

LOGLEVELS = {"NOTSET": logging.NOTSET, "DEBUG": logging.DEBUG, "PROGRESS": logging.INFO - 5, "INFO": logging.INFO,
             "WARNING": logging.WARNING, "ERROR": logging.ERROR, "CRITICAL": logging.CRITICAL}
logging.addLevelName(LOGLEVELS["PROGRESS"], "PROGRESS")


# ----------------------------------------------------------------------------------------------- messages
@dataclass
class AssignLog:
    message: str
    level_str: str
    terminate: bool = False

    @property
    def level_int(self) -> int:
        return LOGLEVELS[self.level_str]


@dataclass
class AssignFile:
    path_audio: str
    dir_results: str
    track: object = None
    duration_audio: float | None = None
    chunklist: list | None = None

    def __post_init__(self):
        self.ident = os.path.splitext(os.path.basename(self.path_audio))[0]
        base = os.path.join(self.dir_results, self.ident)
        self.path_results_partial = base + cfg.SUFFIX_RESULT_PARTIAL
        self.path_results_complete = base + cfg.SUFFIX_RESULT_COMPLETE
        self.shortpath_audio = os.path.basename(self.path_audio)


@dataclass
class AssignChunk:
    file: AssignFile
    chunk: tuple | None = None
    last_chunk: bool = False
    samples: object = None            # pinned PCM view (+ samplerate) handed to the model
    samplerate: int = 16000
    results: object = None
    release: object = None            # callable: hand the pinned buffer back to the streamer's ring


@dataclass
class ExitSignal:
    message: str
    level: str
    end_reason: str


@dataclass
class _Progress:
    outstanding: list = field(default_factory=list)     # chunks streamed but not yet written
    streaming: bool = True                               # the streamer may still add chunks


# ----------------------------------------------------------------------------------------------- coordinator
class Coordinator:
    """Single owner of early / normal exit.  Workers never poll a flag: they call a getter and stop when it returns
    EXIT.  On teardown one EXIT per consumer is put on every queue, so a worker blocked in a bare get() wakes up."""

    def __init__(self, n_gpus: int = 1, inferers_per_gpu: int = 1, streamers_total: int | None = None,
                 depth: int | None = None, q_gui: Queue | None = None, event_analysisdone=None, q_earlyexit=None):
        if n_gpus < 1 or inferers_per_gpu < 1:
            raise ValueError("need at least one GPU and one inferer per GPU")
        self.n_gpus, self.inferers_per_gpu = n_gpus, inferers_per_gpu
        self.analyzers_total = n_gpus * inferers_per_gpu
        # the reference feeds a GPU analyzer from 8 streamers (coordination.py:129-135: resampling on the CPU is the
        # slow part there); here the streamer only reads PCM into pinned memory, 2 per GPU keep a ring full
        self.streamers_total = streamers_total if streamers_total is not None else 2 * n_gpus
        self.queue_depth = depth if depth is not None else max(2, 2 * self.streamers_total // n_gpus)
        self.q_gui = q_gui
        self.q_log: Queue = Queue()
        self.q_stream: Queue = Queue()
        self.q_analyze: list[Queue] = [Queue(maxsize=self.queue_depth) for _ in range(n_gpus)]
        self.q_write: Queue = Queue()
        self.progress: dict[str, _Progress] = {}
        self._lock = threading.Lock()
        self._exit_lock = threading.Lock()
        self.streamers_done, self.analyzers_done, self.writer_done = threading.Event(), threading.Event(), threading.Event()
        self.event_exitanalysis = event_analysisdone if event_analysisdone is not None else threading.Event()
        self.q_earlyexit = q_earlyexit if q_earlyexit is not None else Queue()
        self.end_reason = None

    def log(self, msg: str, level_str: str):
        self.q_log.put(AssignLog(f"coordinator: {msg}", level_str))

    # ---- worker-facing queue API
    def get_stream(self):
        return self.q_stream.get()

    def put_analyze(self, a_chunk: AssignChunk, gpu: int | None = None):
        """Register the chunk as streamed and queue it for a GPU (shortest queue unless `gpu` is given).  Blocks on a
        full queue, but gives up silently once exit has been requested (the streamer finds out on its next get)."""
        with self._lock:
            p = self.progress.setdefault(a_chunk.file.ident, _Progress())
            p.outstanding.append(a_chunk.chunk)
            if a_chunk.last_chunk:
                p.streaming = False
        while not self.event_exitanalysis.is_set():
            g = gpu if gpu is not None else min(range(self.n_gpus), key=lambda i: self.q_analyze[i].qsize())
            try:
                self.q_analyze[g].put(a_chunk, timeout=0.25 if gpu is None else 1.0)
                return
            except Full:
                continue

    def get_analyze(self, gpu: int = 0):
        return self.q_analyze[gpu].get()

    def put_write(self, a_chunk: AssignChunk):
        self.q_write.put(a_chunk)

    def get_write(self):
        """(chunk, fully_analyzed) or EXIT.  A file is fully analyzed when its streamer has finished and every chunk
        it streamed has been written (reference: coordination.py:99-118)."""
        a_chunk = self.q_write.get()
        if a_chunk == EXIT:
            return EXIT
        with self._lock:
            p = self.progress[a_chunk.file.ident]
            p.outstanding.remove(a_chunk.chunk)
            done = not p.outstanding and not p.streaming
        return a_chunk, done

    # ---- exit protocol
    def _poison(self, q: Queue, n: int):
        for _ in range(n):
            q.put(EXIT)

    def exit_analysis(self, sig: ExitSignal):
        """First caller wins (a completion must not overwrite an interruption, nor the reverse).  The logger is NOT
        stopped here: the driver does that after cleanup."""
        with self._exit_lock:
            if self.end_reason is not None:
                return
            self.q_log.put(AssignLog(sig.message, sig.level))
            self.end_reason = sig.end_reason
            self.event_exitanalysis.set()

    def request_stop(self, message: str = "analysis interrupted"):
        self.q_earlyexit.put(message)

    def wait_for_exit(self, threads_streamers, threads_analyzers, thread_writer):
        """threads_analyzers: per GPU, the list of that GPU's inferer threads."""
        def watch_workers():
            for t in threads_streamers:
                t.join()
            self.log("streamers done", "DEBUG")
            self.streamers_done.set()
            for g in range(self.n_gpus):
                self._poison(self.q_analyze[g], self.inferers_per_gpu)
            for per_gpu in threads_analyzers:
                for t in per_gpu:
                    t.join()
            self.log("analyzers done", "DEBUG")
            self.analyzers_done.set()
            self._poison(self.q_write, 1)
            thread_writer.join()
            self.log("writer done", "DEBUG")
            self.writer_done.set()
            self.exit_analysis(ExitSignal("Analysis complete", "INFO", "completed"))

        def watch_queue():
            msg = self.q_earlyexit.get()
            if msg is None:                    # internal wake-up after a normal completion
                return
            self.exit_analysis(ExitSignal(msg, "WARNING", "interrupted"))
            self._poison(self.q_stream, self.streamers_total)
            for g in range(self.n_gpus):
                self._poison(self.q_analyze[g], self.inferers_per_gpu)
            self._poison(self.q_write, 1)

        threading.Thread(target=watch_workers, daemon=True).start()
        threading.Thread(target=watch_queue, daemon=True).start()
        self.event_exitanalysis.wait()
        self.q_earlyexit.put(None)


# ----------------------------------------------------------------------------------------------- manifest lock
FNAME_MANIFEST = "buzzdetect_manifest.json"
KEYS_LOCKED = ("modelname", "output_mode", "classes_out", "precision", "framehop_prop")


def build_manifest(modelname, framehop_prop, precision, classes_out) -> dict:
    """What fixes the schema / resumability of an output folder (reference: manifest.py:13-23)."""
    detections = precision is not None
    return {"modelname": modelname, "output_mode": "detections" if detections else "activations",
            "classes_out": None if detections else sorted(classes_out), "precision": precision,
            "framehop_prop": framehop_prop}


def read_manifest(dir_out: str):
    path = os.path.join(dir_out, FNAME_MANIFEST)
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


def write_manifest(dir_out: str, manifest: dict):
    os.makedirs(dir_out, exist_ok=True)
    with open(os.path.join(dir_out, FNAME_MANIFEST), "w") as f:
        json.dump(manifest, f, indent=2)


def diff_manifests(existing: dict, current: dict) -> list[str]:
    out = []
    for key in KEYS_LOCKED:
        old, new = existing.get(key), current.get(key)
        if key == "classes_out" and old is not None and new is not None:
            so, sn = set(old), set(new)
            if so != sn:
                parts = []
                if sn - so:
                    parts.append("added " + ", ".join(sorted(sn - so)))
                if so - sn:
                    parts.append("removed " + ", ".join(sorted(so - sn)))
                out.append(f"output classes differ ({'; '.join(parts)})")
        elif old != new:
            out.append(f"{key}: existing={old!r}, requested={new!r}")
    return out


def check_or_write_manifest(dir_out: str, manifest: dict):
    """(ok, message): writes the manifest into an empty folder, accepts a matching one, refuses a conflicting one."""
    existing = read_manifest(dir_out)
    if existing is None:
        write_manifest(dir_out, manifest)
        return True, None
    conflicts = diff_manifests(existing, manifest)
    if not conflicts:
        return True, None
    return False, (f"Results have already been written to '{dir_out}' using different settings, so new results would be "
                   "incompatible with the existing files:\n  - " + "\n  - ".join(conflicts) +
                   "\nEither match the existing settings, or choose an empty output folder.")


# ----------------------------------------------------------------------------------------------- log worker
class _Stamp(logging.Formatter):
    def formatTime(self, record, datefmt=None):
        return time.strftime("%Y-%m-%d %H:%M:%S", self.converter(record.created)) + f".{int(record.msecs):03d}"


class WorkerLogger:
    """Drains q_log into `path_log` and the console until a message with terminate=True arrives."""

    def __init__(self, path_log: str, coordinator: Coordinator, verbosity_print: str = "PROGRESS",
                 verbosity_log: str = "DEBUG", log_progress: bool = False, name: str = "buzzdetect_b200.run"):
        self.coordinator = coordinator
        self.level_print = LOGLEVELS[verbosity_print]
        self.log = logging.getLogger(name)
        self.log.setLevel(logging.DEBUG)
        self.log.propagate = False
        fmt = _Stamp("%(asctime)s [%(levelname)s] %(message)s")
        self.h_file = logging.FileHandler(path_log)
        self.h_file.setLevel(LOGLEVELS[verbosity_log])
        if not log_progress:
            self.h_file.addFilter(lambda rec: rec.levelno != LOGLEVELS["PROGRESS"])
        self.h_file.setFormatter(fmt)
        self.h_console = logging.StreamHandler()
        self.h_console.setLevel(self.level_print)
        self.h_console.setFormatter(fmt)
        self.log.addHandler(self.h_file)
        self.log.addHandler(self.h_console)

    def write_log(self, a_log: AssignLog):
        self.log.log(a_log.level_int, a_log.message)
        if self.coordinator.q_gui is not None and a_log.level_int >= self.level_print:
            self.coordinator.q_gui.put(a_log)

    def run(self):
        while True:
            a_log = self.coordinator.q_log.get()
            if a_log.terminate:
                break
            self.write_log(a_log)
        self.write_log(AssignLog("logger closing", "DEBUG"))
        for h in (self.h_file, self.h_console):
            self.log.removeHandler(h)
            h.close()

    __call__ = run


# ----------------------------------------------------------------------------------------------- workers
class StreamWorker:
    """src/stream/worker.py:22-165 around pinned buffers: chunk list (with resume gaps), python-float sample indexing,
    bad reads end the file; the chunk leaves as decoded PCM (downmix + resample happen on the GPU)."""

    def __init__(self, ident, coordinator: Coordinator, chunklength: float, framelength_s: float = 0.96,
                 open_track=None, alloc=None, ring: int = 4):
        from . import pipeline
        self.id, self.coordinator = ident, coordinator
        self.chunklength, self.framelength_s = chunklength, framelength_s
        self.open_track = open_track or pipeline.open_track
        self.alloc = alloc or (lambda shape, dtype: np.empty(shape, dtype))     # capi.pinned_empty on a GPU box
        self.free: Queue = Queue()
        self.ring = ring
        self._bufs = 0

    def log(self, msg, level):
        self.coordinator.q_log.put(AssignLog(f"streamer {self.id}: {msg}", level))

    def _buffer(self, shape, dtype):
        need = int(np.prod(shape))
        while True:
            if self._bufs < self.ring and self.free.empty():
                self._bufs += 1
                return self.alloc(need, dtype)[:need].reshape(shape), need
            b = self.free.get()
            if b.size >= need and b.dtype == np.dtype(dtype):
                return b[:need].reshape(shape), b.size
            self._bufs -= 1                                     # wrong size / type: drop it and allocate a fitting one

    def _chunk_file(self, a_file: AssignFile):
        from . import pipeline
        if os.path.exists(a_file.path_results_complete):
            self.log(f"Skipping {a_file.shortpath_audio}; already analyzed", "DEBUG")
            a_file.chunklist = []
            return
        if os.path.getsize(a_file.path_audio) < pipeline.FILE_SIZE_MINIMUM:
            self.log(f"Skipping {a_file.shortpath_audio}; below minimum analyzeable size", "DEBUG")
            a_file.chunklist = []
            return
        a_file.track = self.open_track(a_file.path_audio)
        a_file.duration_audio = a_file.track.duration
        covered = pipeline._read_partial_starts(a_file.path_results_partial) \
            if os.path.exists(a_file.path_results_partial) else None
        a_file.chunklist = stream.file_chunklist(a_file.duration_audio, self.chunklength, covered, self.framelength_s)
        if covered is not None and covered.size and not a_file.chunklist:
            self.log(f"Discovered non-cleaned file at {a_file.shortpath_audio}; cleaning results", "DEBUG")
            pipeline._finalise(a_file.path_results_partial, a_file.path_results_complete)

    def queue_chunk(self, a_file: AssignFile, chunk, force_last=False) -> bool:
        from . import pipeline
        tr = a_file.track
        sample_from, read_size = stream.chunk_sample_range(chunk, tr.samplerate)
        shape = (read_size, tr.channels) if tr.channels > 1 else (read_size,)
        buf, _ = self._buffer(shape, tr.dtype)
        tr.seek(sample_from)
        n = tr.readinto(buf, read_size)
        cont = n >= read_size
        if not cont:
            pipeline._handle_bad_read(a_file.shortpath_audio, (sample_from + n) / tr.samplerate, a_file.duration_audio, self.log)
            chunk = (chunk[0], round(chunk[0] + (n / tr.samplerate), 1))
        base = buf.base if buf.base is not None else buf
        self.coordinator.put_analyze(AssignChunk(file=a_file, chunk=chunk, samples=buf[:n], samplerate=tr.samplerate,
                                                 last_chunk=force_last or not cont,
                                                 release=lambda b=buf: self.free.put(b.reshape(-1) if b.ndim > 1 else b)))
        del base
        return cont

    def stream_to_queue(self, a_file: AssignFile):
        try:
            self._chunk_file(a_file)
            last = len(a_file.chunklist) - 1
            for i, chunk in enumerate(a_file.chunklist):
                if self.coordinator.event_exitanalysis.is_set():
                    return
                if not self.queue_chunk(a_file, chunk, force_last=i == last):
                    break
        finally:
            if a_file.track is not None:
                a_file.track.close()
                a_file.track = None

    def run(self):
        self.log("launching", "INFO")
        while True:
            a_file = self.coordinator.get_stream()
            if a_file == EXIT:
                break
            self.log(f"buffering {a_file.shortpath_audio}", "INFO")
            self.stream_to_queue(a_file)
        self.log("terminating", "INFO")

    __call__ = run


class InferWorker:
    """src/inference/worker.py:9-92 for one GPU: predict per chunk, hand the (lazy) result to the writer."""

    def __init__(self, ident, gpu: int, model_factory, coordinator: Coordinator):
        self.id, self.gpu, self.coordinator = ident, gpu, coordinator
        self.model_factory = model_factory
        self.model = None

    def log(self, msg, level):
        self.coordinator.q_log.put(AssignLog(f"analyzer {self.id}: {msg}", level))

    def process_chunk(self, a_chunk: AssignChunk):
        t0 = time.perf_counter()
        if hasattr(self.model, "predict_pcm"):
            a_chunk.results = self.model.predict_pcm(a_chunk.samples, a_chunk.samplerate)
        else:
            a_chunk.results = self.model.predict(a_chunk.samples)
        self.coordinator.put_write(a_chunk)
        dur = a_chunk.chunk[1] - a_chunk.chunk[0]
        dt = max(time.perf_counter() - t0, 1e-9)
        self.log(f"queued {a_chunk.file.shortpath_audio}, chunk ({float(a_chunk.chunk[0]):.2f}, {float(a_chunk.chunk[1]):.2f}) "
                 f"in {dt:.4f}s (rate: {dur / dt:.1f})", "PROGRESS")

    def run(self):
        self.log(f"launching on GPU {self.gpu}", "INFO")
        self.model = self.model_factory(self.gpu)
        while True:
            a_chunk = self.coordinator.get_analyze(self.gpu)
            if a_chunk == EXIT:
                break
            self.process_chunk(a_chunk)
        self.log("terminating", "DEBUG")

    __call__ = run


class WriteWorker:
    """src/write/worker.py:10-100 without pandas: results.numpy() -> rows -> append; sort + promote when the file is
    fully analyzed."""

    def __init__(self, classes, framehop_s, dir_out, coordinator: Coordinator, threshold=None, classes_out="all",
                 digits_time: int = 2, digits_results: int = 2):
        self.classes, self.framehop_s, self.dir_out, self.coordinator = classes, framehop_s, dir_out, coordinator
        self.threshold, self.classes_out = threshold, classes_out
        self.digits_time, self.digits_results = digits_time, digits_results
        self.frames = 0

    def log(self, msg, level):
        self.coordinator.q_log.put(AssignLog(f"writer: {msg}", level))

    def write_results(self, a_chunk: AssignChunk, fully_analyzed: bool):
        from . import pipeline
        act = np.asarray(a_chunk.results.numpy())
        if a_chunk.release is not None:
            a_chunk.release()                                   # the pinned buffer goes back to the streamer's ring
            a_chunk.samples = None
        if self.threshold is None:
            cols, start, vals = write.format_activations(act, self.classes, self.framehop_s, self.digits_time,
                                                         time_start=a_chunk.chunk[0], classes_keep=self.classes_out,
                                                         digits_results=self.digits_results)
        else:
            cols, start, vals = write.format_detections(act, self.threshold, self.classes, self.framehop_s,
                                                        self.digits_time, a_chunk.chunk[0])
        os.makedirs(os.path.dirname(a_chunk.file.path_results_partial) or ".", exist_ok=True)
        pipeline._append_rows(a_chunk.file.path_results_partial, cols, start, vals)
        self.frames += act.shape[0]
        if fully_analyzed:
            pipeline._finalise(a_chunk.file.path_results_partial, a_chunk.file.path_results_complete)

    def run(self):
        self.log("launching", "INFO")
        while True:
            item = self.coordinator.get_write()
            if item == EXIT:
                break
            self.write_results(*item)
        self.log("terminating", "DEBUG")

    __call__ = run


# ----------------------------------------------------------------------------------------------- driver
def run_analysis(paths: list[str], dir_out: str, gpus: list[int] | None = None, modelname: str = cfg.DEFAULT_MODEL,
                 chunklength: float = 199.68, framehop_prop: float = 1.0, precision: float | None = None,
                 classes_out="all", model_factory=None, open_track=None, alloc=None, streamers: int | None = None,
                 verbosity_print: str = "WARNING", stop_after: float | None = None) -> dict:
    """Files -> result CSVs on `gpus` (default: GPU 0) in one process.  `precision`: None writes activations, a number
    writes detections at the threshold that gives this precision (write.calculate_threshold).  Returns a report."""
    from . import pipeline
    gpus = list(gpus) if gpus else [0]
    os.makedirs(dir_out, exist_ok=True)
    if model_factory is None:
        from .inference.models import load_model

        factory_lock = threading.Lock()

        def model_factory(gpu):
            with factory_lock:                      # the plugin reads its device from the environment at initialize()
                os.environ["BUZZ_B200_DEVICE"] = str(gpu)
                return load_model(modelname, framehop_prop=framehop_prop, initialize=True)
        if alloc is None:
            from . import capi
            alloc = capi.pinned_empty
    probe = model_factory.__dict__.get("describe") if hasattr(model_factory, "__dict__") else None
    # attributes the streamer / writer need BEFORE any model is initialised
    if probe is not None:
        classes, framelength_s = probe["classes"], probe.get("framelength_s", 0.96)
    else:
        from .inference.models import load_model
        m0 = load_model(modelname, framehop_prop=framehop_prop, initialize=False)
        classes, framelength_s = m0.config["classes"], m0.embedder.framelength_s
    pipeline.validate_framehop(framehop_prop)
    manifest = build_manifest(modelname, framehop_prop, precision, classes if classes_out == "all" else classes_out)
    ok, msg = check_or_write_manifest(dir_out, manifest)
    if not ok:
        raise ValueError(msg)
    threshold = write.calculate_threshold(modelname, precision) if precision is not None else None
    chunklength = stream.setup_chunklength(chunklength, framelength_s)
    co = Coordinator(n_gpus=len(gpus), streamers_total=streamers)
    logger = WorkerLogger(os.path.join(dir_out, "buzzdetect_b200.log"), co, verbosity_print=verbosity_print)
    t_log = threading.Thread(target=logger, daemon=True)
    t_log.start()
    for p in paths:
        co.q_stream.put(AssignFile(path_audio=p, dir_results=dir_out))
    co._poison(co.q_stream, co.streamers_total)                # streamers stop when the file list is exhausted
    t_stream = [threading.Thread(target=StreamWorker(i, co, chunklength, framelength_s, open_track, alloc), daemon=True)
                for i in range(co.streamers_total)]
    t_infer = [[threading.Thread(target=InferWorker(f"{g}", gi, lambda _gi, g=g: model_factory(g), co), daemon=True)]
               for gi, g in enumerate(gpus)]
    writer = WriteWorker(classes, framelength_s * framehop_prop, dir_out, co, threshold=threshold, classes_out=classes_out)
    t_write = threading.Thread(target=writer, daemon=True)
    t0 = time.perf_counter()
    for t in t_stream + [t for per in t_infer for t in per] + [t_write]:
        t.start()
    if stop_after is not None:
        threading.Timer(stop_after, co.request_stop, args=("stop requested",)).start()
    co.wait_for_exit(t_stream, t_infer, t_write)
    dt = time.perf_counter() - t0
    co.q_log.put(AssignLog("", "DEBUG", terminate=True))
    t_log.join(timeout=5)
    return {"end_reason": co.end_reason, "seconds": dt, "frames": writer.frames, "gpus": gpus}
