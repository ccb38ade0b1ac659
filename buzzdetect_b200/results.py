"""Objects handed across the reference's queues.

The reference stores a tf.Tensor in AssignChunk.results and the writer calls ``results.numpy()``
(src/write/worker.py:69).  DeviceResults offers the same ``.numpy()`` (plus ``__array__``) over a host buffer that
the CUDA stream has finished writing by the time numpy() returns, so it is safe to move to another thread."""
import numpy as np


class DeviceResults:
    """Results of one chunk.  Built either over a finished host array or over a capi.Ticket (chunk still in flight:
    predict() has returned, the GPU may not have); numpy() / __array__ block until the activations are on the host."""

    __slots__ = ("_host", "_ticket", "_which", "embeddings")

    def __init__(self, host=None, embeddings: np.ndarray | None = None, ticket=None, which: str = "act"):
        self._host = host
        self._ticket = ticket
        self._which = which
        self.embeddings = embeddings

    def numpy(self) -> np.ndarray:
        if self._host is None:
            tk = self._ticket.wait()
            self._host = tk.act if self._which == "act" else tk.emb
            self._ticket = None
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, k):
        return self.numpy()[k]

    def __len__(self):
        return len(self.numpy())

    @property
    def shape(self):
        return self.numpy().shape

    @property
    def dtype(self):
        return np.dtype(np.float32)


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
