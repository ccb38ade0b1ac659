"""The C-ABI library loads on a CPU-only box and exports every symbol include/buzzdetect_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "buzzdetect_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bd_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(built_lib):
    from buzzdetect_b200 import capi
    syms = header_symbols()
    assert len(syms) >= 15
    assert sorted(capi.SIGNATURES) == syms
    for s in syms:
        assert getattr(built_lib, s) is not None


def test_library_has_no_torch_dependency(built_lib):
    import subprocess
    from buzzdetect_b200 import capi
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "tensorflow" not in out


def test_sass_contains_blackwell_paths(built_lib):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md)."""
    import shutil
    import subprocess
    from buzzdetect_b200 import capi
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnem in sass, mnem
    assert "arch = sm_100a" in sass


def test_engine_create_fails_loudly_without_gpu(built_lib, yamnet_variables):
    """No CPU fallback: without a B200 the product path must raise, never route through the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from buzzdetect_b200 import capi
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        capi.Engine(device=0, yamnet_variables=yamnet_variables)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "buzzdetect_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)
