import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# the YAMNet blob is not in the reference checkout: tests run on the seeded synthetic network (explicit opt-in)
os.environ.setdefault("BUZZ_B200_ALLOW_SYNTHETIC", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, built in-tree (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    g.build()
    from buzzdetect_b200 import capi
    return capi.load_library()


@pytest.fixture(scope="session")
def yamnet_variables():
    from buzzdetect_b200 import weights as W
    v, prov = W.resolve_yamnet(verify=True)
    return v


@pytest.fixture(scope="session")
def head():
    from buzzdetect_b200 import weights as W
    return W.load_head()


@pytest.fixture(scope="session")
def mel():
    from buzzdetect_b200 import weights as W
    return W.load_mel()


@pytest.fixture(scope="session")
def engines(built_lib, yamnet_variables):
    """One engine per precision on cuda:0, shared by the GPU tests (small sub-batches to exercise the loops)."""
    from buzzdetect_b200 import capi
    cache = {}

    def get(precision="fp16x3", **kw):
        key = (precision, tuple(sorted(kw.items())))
        if key not in cache:
            cache[key] = capi.Engine(device=0, yamnet_variables=yamnet_variables, precision=precision, **kw)
        return cache[key]

    yield get
    for e in cache.values():
        e.close()


_REPORT = {}


def report(key, val):
    """Append a measured number to gpurun_out/parity_report.json (read back into profiles/ and DESIGN.md)."""
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_report.json")
        if not _REPORT and os.path.exists(path):
            with open(path) as f:
                _REPORT.update(json.load(f))          # several pytest invocations append to one report
        _REPORT[key] = val
        with open(path, "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.fixture(scope="session")
def parity_report():
    return report
