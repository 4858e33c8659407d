"""The drop-in claim of INTEGRATION.md §2, run: oracle/_ref/ref_gpu is the REFERENCE program — its own stdin protocol,
XML / OBJ / MTL loaders, buildBVH, imshow and svpng, compiled unmodified — with only the sample loop main.cpp:79-113
replaced (integration/main_gpu.sed) by integration/trt_bridge.cpp + trt_render.  Its PNG must equal the PNG of this
repo's own host mirror (bin/trt_main) for the same seed: same scene files in, same picture out, whichever host side
feeds the library."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, SCENES

pytestmark = pytest.mark.gpu
REF_GPU = os.path.join(ROOT, "oracle", "_ref", "ref_gpu")
TRT_MAIN = os.path.join(ROOT, "tinyraytracing_b200", "bin", "trt_main")


def _run(exe, f, spp, outdir, seed):
    import cv2

    os.makedirs(outdir, exist_ok=True)
    # the programs write <basedir>/image<SPP>.png: give each its own base directory with the scene's files linked in
    for fn in os.listdir(f["basedir"]):
        dst = os.path.join(outdir, fn)
        if not os.path.exists(dst):
            os.symlink(os.path.join(f["basedir"], fn), dst)
    rel = lambda p: os.path.join(outdir, os.path.relpath(p, f["basedir"]))
    inp = "%s\n%s\n%s\n%s\n%d\n" % (outdir, rel(f["mtl"]), rel(f["xml"]), rel(f["obj"]), spp)
    env = dict(os.environ, TRT_SEED=str(seed))
    r = subprocess.run([exe], input=inp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, env=env)
    assert r.returncode == 0, (exe, r.stdout[-400:], r.stderr[-400:])
    png = os.path.join(outdir, "image%d.png" % spp)
    assert os.path.exists(png), r.stdout[-400:]
    return cv2.imread(png, cv2.IMREAD_COLOR)


@pytest.mark.parametrize("name", SCENES)
def test_reference_program_with_bridge_renders_what_trt_main_renders(name, scene_files, tmp_path):
    if not os.path.exists(REF_GPU):
        pytest.skip("oracle/_ref/ref_gpu not built (make -C oracle bridge needs /root/reference: build container only)")
    f = scene_files[name]
    a = _run(REF_GPU, f, 6, str(tmp_path / "ref_gpu"), seed=31)
    b = _run(TRT_MAIN, f, 6, str(tmp_path / "trt_main"), seed=31)
    assert a is not None and b is not None and a.shape == b.shape
    assert a.mean() > 1  # a picture, not a black frame
    assert np.array_equal(a, b), (name, float((a != b).mean()))
