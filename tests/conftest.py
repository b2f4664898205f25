import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def native_libs():
    """Build the in-tree shared libraries when they are missing (nvcc cross-compiles without a GPU)."""
    from cnosdb_b200 import build, cabi
    need = [cabi.gpu_library_path(), cabi.hostgen_library_path(), os.path.join(ROOT, "oracle", "libtskv_oracle.so")]
    if not all(os.path.exists(p) for p in need[:2]):
        build.build_all()
    if not os.path.exists(need[2]):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return need


@pytest.fixture(scope="session")
def golden():
    import json
    d = os.path.join(ROOT, "tests", "golden")
    out = {}
    for name in ("codec_vectors", "window_kat", "sql_goldens"):
        with open(os.path.join(d, name + ".json")) as f:
            out[name] = json.load(f)
    return out


@pytest.fixture(scope="session")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from cnosdb_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()
