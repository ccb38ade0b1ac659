"""Objects handed across the reference's queues.

The reference stores a tf.Tensor in AssignChunk.results and the writer calls ``results.numpy()``
(src/write/worker.py:69).  DeviceResults offers the same ``.numpy()`` (plus ``__array__``) over a host buffer that
the CUDA stream has already finished writing, so it is safe to move to another thread."""
import numpy as np


class DeviceResults:
    __slots__ = ("_host", "embeddings")

    def __init__(self, host: np.ndarray, embeddings: np.ndarray | None = None):
        self._host = host
        self.embeddings = embeddings

    def numpy(self) -> np.ndarray:
        return self._host

    def __array__(self, dtype=None, copy=None):
        return self._host if dtype is None else self._host.astype(dtype)

    def __getitem__(self, k):
        return self._host[k]

    def __len__(self):
        return len(self._host)

    @property
    def shape(self):
        return self._host.shape

    @property
    def dtype(self):
        return self._host.dtype


def as_host_f32(samples) -> np.ndarray:
    """Accept what the streamer may hand over: numpy, a CPU torch tensor (zero-copy), or anything array-like."""
    if isinstance(samples, np.ndarray):
        a = samples
    elif hasattr(samples, "detach") and hasattr(samples, "cpu"):        # torch.Tensor
        a = samples.detach().cpu().numpy()
    elif hasattr(samples, "numpy"):
        a = samples.numpy()
    else:
        a = np.asarray(samples)
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 1:
        raise ValueError(f"expected 1-D mono samples, got shape {a.shape}")
    return a
