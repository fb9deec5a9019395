import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    # -m gpu tests need the device; skip them cleanly (not fail) if someone runs the whole suite on CPU
    import importlib
    try:
        n = importlib.import_module("julia-raytracer_b200._lib").lib().jt_device_count()
    except Exception:
        n = 0
    if n > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build()
    yield


SCENE_DIR = os.path.join(ROOT, "assets", "scenes")


@pytest.fixture(scope="session")
def scenes():
    """name -> (SceneData, SceneBvh, lights), loaded lazily from the packed assets."""
    import importlib
    jt = importlib.import_module("julia-raytracer_b200")
    bvh = importlib.import_module("julia-raytracer_b200.bvh")
    lights = importlib.import_module("julia-raytracer_b200.lights")
    cache = {}

    def get(name):
        if name not in cache:
            if name.startswith("synthetic"):
                import synth
                sc = synth.make_scene(name)
            else:
                sc = jt.load_scene(os.path.join(SCENE_DIR, f"{name}.jtscene"))
            cache[name] = (sc, bvh.make_scene_bvh(sc), lights.make_trace_lights(sc))
        return cache[name]

    return get
