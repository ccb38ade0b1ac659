"""Boundary shim, not built work: the reference's plugin interface src/inference/embedding.py:8-79 (BaseEmbedder, load_embedder) kept VERBATIM in names,
attributes, argument meaning and discovery rules, so that the plugins behave identically when they are loaded by the
reference's own module instead (inside a buzzdetect checkout they import `src.inference.*` first; see
tests/test_reference_interop.py::test_plugins_load_through_the_reference_loader)."""
import importlib.util
from abc import ABC, abstractmethod
from pathlib import Path

from buzzdetect_b200 import config as cfg


class BaseEmbedder(ABC):
    """Abstract base class for all audio embedders (reference: embedding.py:8-37)."""

    embeddername: str = None
    samplerate: int = None
    framelength_s: float = None
    n_embeddings: int = None
    digits_time: int = None
    dtype_in: str = None

    def __init__(self, framehop_prop):
        # cheap and GPU-free: Analyzer / WorkerStreamer read these attributes before initialize()
        self.framehop_prop = framehop_prop
        self.framehop_s = self.framelength_s * framehop_prop
        self.model = None

    @abstractmethod
    def initialize(self):
        pass

    @abstractmethod
    def embed(self, samples):
        pass


def load_embedder(embeddername: str, framehop_prop: float, initialize: bool):
    """reference: embedding.py:40-79 -- import <DIR_EMBEDDERS>/<name>/embedder.py, first BaseEmbedder subclass wins."""
    embedder_path = Path(cfg.DIR_EMBEDDERS) / embeddername
    if not embedder_path.exists():
        raise ValueError(f"Embedder '{embeddername}' not found in {cfg.DIR_EMBEDDERS}")

    spec = importlib.util.spec_from_file_location(f"{embeddername}_embedder", embedder_path / "embedder.py")
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)

    embedder_class = None
    for item_name in dir(module):
        item = getattr(module, item_name)
        if isinstance(item, type) and issubclass(item, BaseEmbedder) and item is not BaseEmbedder \
                and item.__module__ == module.__name__:
            embedder_class = item
            break
    if embedder_class is None:
        raise ValueError(f"No BaseEmbedder subclass found in {embeddername}/embedder.py")

    embedder = embedder_class(framehop_prop=framehop_prop)
    if initialize:
        embedder.initialize()
    return embedder
