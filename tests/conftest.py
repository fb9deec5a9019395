import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")
    # build BEFORE the device probe below: on a fresh checkout the probe would otherwise fail to load the library,
    # every gpu test would be skipped and `pytest -m gpu` would exit 0 without having run a single parity test
    import __graft_entry__ as g
    g.build()


def pytest_collection_modifyitems(config, items):
    import importlib
    n = importlib.import_module("julia-raytracer_b200._lib").lib().jt_device_count()
    if n > 0:
        return
    # `-m gpu` asked for explicitly on a machine without a device is an error, not a green run of zero tests
    markexpr = (config.getoption("-m") or "").strip()
    if markexpr == "gpu":
        raise pytest.UsageError("-m gpu selected but libjtrace_b200 sees no CUDA device (there is no CPU fallback)")
    # the whole suite on a CPU box: gpu tests are skipped, everything else runs
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
