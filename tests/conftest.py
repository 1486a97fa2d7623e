import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_module(name):
    """Import tests/golden/<name>.py (the fixture generators double as helper libraries for the tests)."""
    import importlib.util
    key = "golden_" + name
    if key not in sys.modules:
        spec = importlib.util.spec_from_file_location(key, os.path.join(GOLDEN, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[key] = mod
        spec.loader.exec_module(mod)
    return sys.modules[key]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a host where the library finds no CUDA device — a plain `pytest` on a
    CPU box then reports skips instead of 'no CUDA device visible' errors."""
    if not any("gpu" in it.keywords for it in items):
        return
    try:
        from tredparse_b200 import _lib
        have = _lib.load().tredsw_device_count() > 0
    except Exception:
        return                       # the library itself is missing or broken: let the tests fail loudly
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (run on the B200 box with -m gpu)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (gcc) once per session; the CUDA library only when nvcc is around and the
    in-tree .so is stale or absent (on the GPU box the .so shipped with the snapshot is used)."""
    from oracle import sw
    sw.build()
    from tredparse_b200 import build as b
    try:
        b.build()
    except Exception as e:  # pragma: no cover
        if not os.path.exists(b.OUT):
            raise
        sys.stderr.write("libtredsw.so rebuild skipped: {}\n".format(e))
    yield
