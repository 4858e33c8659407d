"""The drop-in driver (tinyraytracing_b200/bin/trt_main: the reference's stdin protocol, main.cpp:46-55, with the
sample loop replaced by the GPU path) and the resolve / 8-bit pack of imshow (main.cpp:30-38)."""
import os
import subprocess

import numpy as np
import pytest

import tinyraytracing_b200 as trt

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gamma_pack(img):
    # (unsigned char)clamp(pow(v, 1.0f / 2.2f) * 255, 0.0, 255.0)
    g = np.power(img, np.float64(np.float32(1.0) / np.float32(2.2))) * 255
    return np.minimum(np.maximum(g, 0.0), 255.0).astype(np.uint8)


def test_resolve_pack_matches_imshow(device_scenes):
    import torch

    dev = device_scenes["back"]
    acc = torch.zeros(dev.height * dev.width * 3, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    dev.render_accumulate(dev.params(4, seed=3), acc.data_ptr())
    img, rgb = dev.resolve(acc.data_ptr(), 4, want_rgb8=True)
    assert np.array_equal(img, dev.render(4, seed=3))
    ref = gamma_pack(img)
    # device pow vs libm pow may differ in the last ulp: at most one 8-bit step, on a vanishing fraction of pixels
    diff = np.abs(rgb.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3


def test_trt_main_renders_the_scene_files(scene_files, device_scenes, tmp_path):
    import cv2

    exe = os.path.join(ROOT, "tinyraytracing_b200", "bin", "trt_main")
    assert os.path.exists(exe), "bin/trt_main not built"
    f = scene_files["back"]
    spp = 4
    inp = "%s\n%s\n%s\n%s\n%d\n" % (f["basedir"], f["mtl"], f["xml"], f["obj"], spp)
    env = dict(os.environ, TRT_SEED="17")
    r = subprocess.run([exe], input=inp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Build BVH down." in r.stdout and "Iamge output to" in r.stdout  # the reference's own progress lines
    png = cv2.imread(os.path.join(f["basedir"], "image%d.png" % spp), cv2.IMREAD_COLOR)[:, :, ::-1]
    img = device_scenes["back"].render(spp, seed=17)
    assert png.shape == img.shape
    assert np.array_equal(png, gamma_pack(img))


@pytest.mark.parametrize("name", ("back", "staircase"))
def test_cpp_host_surface_selftest(name, scene_files):
    """csrc/host/selftest.cpp: the C++ classes and functions a reference user calls (Scene loaders, buildBVH,
    traverseBVH -> HitRecord, the render loop) against the GPU-backed implementations."""
    exe = os.path.join(ROOT, "tinyraytracing_b200", "bin", "trt_host_selftest")
    assert os.path.exists(exe), "bin/trt_host_selftest not built"
    f = scene_files[name]
    r = subprocess.run([exe, f["basedir"], f["mtl"], f["xml"], f["obj"]], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=300)
    assert r.returncode == 0 and "host selftest ok" in r.stdout, r.stdout[-2000:]


def _trt_main(f, spp, env_extra, outdir):
    import cv2

    os.makedirs(outdir, exist_ok=True)
    for fn in os.listdir(f["basedir"]):
        dst = os.path.join(outdir, fn)
        if not os.path.exists(dst):
            os.symlink(os.path.join(f["basedir"], fn), dst)
    rel = lambda p: os.path.join(outdir, os.path.relpath(p, f["basedir"]))
    inp = "%s\n%s\n%s\n%s\n%d\n" % (outdir, rel(f["mtl"]), rel(f["xml"]), rel(f["obj"]), spp)
    exe = os.path.join(ROOT, "tinyraytracing_b200", "bin", "trt_main")
    r = subprocess.run([exe], input=inp, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300,
                       env=dict(os.environ, TRT_SEED="23", **env_extra))
    assert r.returncode == 0, r.stdout[-600:] + r.stderr[-600:]
    return cv2.imread(os.path.join(outdir, "image%d.png" % spp), cv2.IMREAD_COLOR), r.stdout


def test_trt_main_multi_device_and_checkpoint(scene_files, tmp_path):
    """The C++ driver's round-2 switches: TRT_DEVICES (here the same GPU twice with the library's peer reduce, which a
    one-GPU box can run; the NCCL flavour on distinct GPUs is tests/test_gpu_entrypoints.py's) and TRT_CHECKPOINT
    (rendered in steps, then resumed from the finished file) give the PNG of the plain run."""
    f = scene_files["veach-mis"]
    plain, _ = _trt_main(f, 7, {}, str(tmp_path / "plain"))
    multi, out = _trt_main(f, 7, {"TRT_DEVICES": "0,0", "TRT_PEER_REDUCE": "1"}, str(tmp_path / "multi"))
    assert "rendered on 2 GPUs (peer-memory reduce)" in out
    assert np.array_equal(plain, multi)
    ck = str(tmp_path / "ck")
    step, out = _trt_main(f, 7, {"TRT_CHECKPOINT": os.path.join(ck, "frame.ckpt"), "TRT_CHECKPOINT_EVERY": "3"}, ck)
    assert "7 of 7 samples rendered by this run" in out and np.array_equal(plain, step)
    again, out = _trt_main(f, 7, {"TRT_CHECKPOINT": os.path.join(ck, "frame.ckpt"), "TRT_CHECKPOINT_EVERY": "3"}, ck)
    assert "0 of 7 samples rendered by this run" in out and np.array_equal(plain, again)
