import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "scripts")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def rt():
    """The product package (ctypes mirror of librt_b200.so); builds the library if missing."""
    import rtb200
    if not os.path.exists(rtb200.library_path()):
        rtb200.build_library()
    rtb200.load_library()
    return rtb200


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    oracle_py.lib()
    return oracle_py


def scene_path(name):
    return os.path.join(GOLDEN, "scenes", name + ".txt")


@pytest.fixture(scope="session")
def scenes(rt):
    return {n: rt.load_scene(scene_path(n)) for n in ("simple", "medium", "complex")}
