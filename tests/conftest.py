import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENES = os.path.join(ROOT, "scenes")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices():
    try:
        import raytracing_course_b200 as m
        m.load_library()
        return m.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device skips the gpu-marked tests instead of failing them
    (the product has no CPU path to fall back to)."""
    if not any("gpu" in it.keywords for it in items):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: the product has no CPU path")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def scene_path(name):
    """Path of scenes/<name>.txt, generating the scene files on first use."""
    p = os.path.join(SCENES, name + ".txt")
    if not os.path.exists(p):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "make_scenes.py")])
    return p


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def oracle_lib():
    import orclib
    return orclib.oracle()


@pytest.fixture(scope="session")
def rtc():
    """The product binding.  The library must already be built (python -c 'import
    __graft_entry__ as g; g.build()'); tests never build or fall back silently."""
    import raytracing_course_b200 as m
    m.load_library()
    return m


@pytest.fixture(scope="session")
def gpu_scenes(rtc):
    """Scenes resident on cuda:0, cached for the session."""
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = rtc.Scene(path=scene_path(name), device=0)
        return cache[name]

    yield get
    for s in cache.values():
        s.close()


@pytest.fixture(scope="session")
def oracle_scenes(oracle_lib):
    import orclib
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = orclib.Scene(oracle_lib, scene_path(name))
        return cache[name]

    yield get
    for s in cache.values():
        s.close()
