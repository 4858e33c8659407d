import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


SCENES = ("back", "veach-mis", "staircase")
# reduced resolutions (same aspect) used by the render parity tests: the reference takes W,H from the XML
SMALL_RES = {"back": (64, 64), "veach-mis": (96, 54), "staircase": (96, 54)}


@pytest.fixture(scope="session")
def scene_files(tmp_path_factory):
    """Materialised OBJ/MTL/XML (+ pre-decoded textures) of the packed cg22 scenes, at reduced resolution."""
    from tinyraytracing_b200 import scenes

    out = {}
    for name in SCENES:
        d = tmp_path_factory.mktemp("scn_" + name.replace("-", "_"))
        w, h = SMALL_RES[name]
        out[name] = scenes.materialize(name, str(d), width=w, height=h)
    return out


@pytest.fixture(scope="session")
def host_scenes(scene_files):
    import tinyraytracing_b200 as trt

    return {n: trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"]) for n, f in scene_files.items()}


@pytest.fixture(scope="session")
def oracle_scenes():
    import oraclelib

    return {n: oraclelib.OracleScene(oraclelib.parsed_scene(n, *SMALL_RES[n])) for n in SCENES}


@pytest.fixture(scope="session")
def device_scenes(host_scenes):
    import tinyraytracing_b200 as trt

    devs = {n: trt.DeviceScene(h, 0) for n, h in host_scenes.items()}
    yield devs
    for d in devs.values():
        d.close()


def make_rays(host, oracle, n, seed):
    """The BASELINE config-2 ray population, surface points supplied by the ORACLE (CPU)."""
    from tinyraytracing_b200 import workloads

    def tracer(rays):
        ids, t, pn, hp = oracle.trace(rays, want_pn=True)
        return ids, hp, pn

    return workloads.fixed_ray_batch(n, host.camera(), host.root_box(), tracer, seed=seed)
