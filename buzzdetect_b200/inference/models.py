"""Boundary shim, not built work: the reference's plugin interface src/inference/models.py:12-79 (BaseModel, load_model) kept VERBATIM in names,
attributes, argument meaning and discovery rules, so that the plugins behave identically when they are loaded by the
reference's own module instead (inside a buzzdetect checkout they import `src.inference.*` first; see
tests/test_reference_interop.py::test_plugins_load_through_the_reference_loader)."""
import importlib.util
import json
import os
from abc import ABC, abstractmethod
from pathlib import Path

from buzzdetect_b200 import config as cfg
from buzzdetect_b200.inference.embedding import BaseEmbedder, load_embedder


class BaseModel(ABC):
    """Abstract base class for all buzzdetect models (reference: models.py:12-37)."""

    modelname: str = None
    embeddername: str = None
    digits_results: int = None
    dtype_in: str = None

    def __init__(self, framehop_prop):
        self.model = None
        self.embedder: BaseEmbedder = load_embedder(embeddername=self.embeddername, framehop_prop=framehop_prop,
                                                    initialize=False)
        with open(os.path.join(cfg.DIR_MODELS, self.modelname, "config_model.json"), "r") as f:
            self.config = json.load(f)

    @abstractmethod
    def initialize(self):
        pass

    @abstractmethod
    def predict(self, audiosamples):
        pass


def load_model(modelname: str, framehop_prop: float, initialize: bool):
    """reference: models.py:40-79."""
    model_path = Path(cfg.DIR_MODELS) / modelname
    if not model_path.exists():
        raise ValueError(f"model '{modelname}' not found in {cfg.DIR_MODELS}")

    spec = importlib.util.spec_from_file_location(f"{modelname}_model", model_path / "model.py")
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)

    model_class = None
    for item_name in dir(module):
        item = getattr(module, item_name)
        if isinstance(item, type) and issubclass(item, BaseModel) and item is not BaseModel \
                and item.__module__ == module.__name__:
            model_class = item
            break
    if model_class is None:
        raise ValueError(f"No BaseModel subclass found in {modelname}/model.py")

    model = model_class(framehop_prop=framehop_prop)
    if initialize:
        model.initialize()
    return model
