"""The C++ baseline-JPEG decoder behind Material::readinMap (csrc/host/jpeg_decoder.cpp) must return exactly the
bytes cv::imread returns (the reference decodes textures with OpenCV, material.cpp:6, and texels feed Kd directly):
compared against cv2 on JPEGs of every chroma sampling the decoder accepts, grayscale, restart intervals, optimised
Huffman tables, odd and tiny sizes, and on the three packed cg22 textures.  CPU only."""
import json
import os

import cv2
import numpy as np
import pytest

import tinyraytracing_b200 as trt
from tinyraytracing_b200 import api, scenes


def synth(h, w, seed, gray=False):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(xx / 7.0 + k) * np.cos(yy / (5.0 + k)) for k in range(3)], -1)
    img += rng.normal(0, 12, img.shape)
    img[h // 3: h // 2, w // 4: w // 2] = (250, 10, 128)  # a hard edge: exercises the chroma filters' rounding
    img = np.clip(img, 0, 255).astype(np.uint8)
    return img[:, :, 0].copy() if gray else img


SAMPLINGS = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
             "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440}


@pytest.mark.parametrize("sampling", list(SAMPLINGS))
@pytest.mark.parametrize("size", [(64, 64), (37, 53), (1, 1), (2, 3), (17, 4), (8, 200), (131, 129)])
def test_decoder_matches_opencv(tmp_path, sampling, size):
    h, w = size
    img = synth(h, w, seed=h * 1000 + w)
    for quality, extra in ((90, []), (35, [cv2.IMWRITE_JPEG_OPTIMIZE, 1]), (75, [cv2.IMWRITE_JPEG_RST_INTERVAL, 3])):
        ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, quality, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
                                             SAMPLINGS[sampling]] + extra)
        assert ok
        p = str(tmp_path / "t.jpg")
        enc.tofile(p)
        ref = cv2.imread(p, cv2.IMREAD_COLOR)
        mine = api.decode_jpeg(p)
        assert mine.shape == ref.shape
        assert np.array_equal(mine, ref), (sampling, size, quality, int(np.abs(mine.astype(int) - ref.astype(int)).max()))


def test_grayscale_and_errors(tmp_path):
    g = synth(45, 70, 3, gray=True)
    p = str(tmp_path / "g.jpg")
    cv2.imencode(".jpg", g, [cv2.IMWRITE_JPEG_QUALITY, 80])[1].tofile(p)
    assert np.array_equal(api.decode_jpeg(p), cv2.imread(p, cv2.IMREAD_COLOR))
    cv2.imencode(".jpg", synth(40, 40, 4), [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])[1].tofile(p)
    with pytest.raises(trt.TrtError) as e:
        api.decode_jpeg(p)
    assert "progressive" in str(e.value)
    with open(p, "wb") as f:
        f.write(b"not a jpeg at all")
    with pytest.raises(trt.TrtError):
        api.decode_jpeg(p)
    with pytest.raises(trt.TrtError):
        api.decode_jpeg(str(tmp_path / "missing.jpg"))


def test_cg22_textures_decode_identically(tmp_path):
    z = np.load(os.path.join(scenes.SCENE_DIR, "staircase.npz"))
    meta = json.loads(str(z["meta"]))
    assert len(meta["textures"]) == 3
    for rel in meta["textures"]:
        p = str(tmp_path / os.path.basename(rel))
        z["jpeg:" + rel].tofile(p)
        assert np.array_equal(api.decode_jpeg(p), cv2.imread(p, cv2.IMREAD_COLOR)), rel


def test_loader_decodes_textures_without_sidecar(tmp_path):
    """Material::readinMap decodes the JPEG itself: remove the pre-decoded side-cars and load staircase."""
    f = scenes.materialize("staircase", str(tmp_path), width=32, height=18)
    ref = {tuple(t.shape): t for t in trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"]).textures()}
    for root, _, files in os.walk(str(tmp_path)):
        for name in files:
            if name.endswith(".bgr"):
                os.remove(os.path.join(root, name))
    tex = trt.HostScene.load(f["xml"], f["obj"], f["mtl"], f["basedir"]).textures()
    assert len(tex) == 3
    for t in tex:
        assert np.array_equal(t, ref[tuple(t.shape)])
        assert t.any()
