"""Round-2 entry points of include/trt.h through ctypes: trt_shade (pathtracing.h:14 as a batch), trt_accum_* (checkpoint /
resume), trt_render_multi (the multi-GPU fan-out inside the library, main.cpp:79-81), trt_stats.rays_strict."""
import os

import numpy as np
import pytest

from conftest import SCENES, make_rays

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tail", ("default", "0"))  # k_finish from the first vertex (6000 records) / per-depth kernels only
@pytest.mark.parametrize("name", SCENES)
def test_shade_batch_matches_oracle(name, tail, host_scenes, oracle_scenes, device_scenes, monkeypatch):
    """shade(hit, wi) for a batch of traced rays: NEE + Russian roulette + the whole bounce chain from caller-supplied
    hits, against the oracle's recursive shade() on the same records with the same Philox streams (pixel key = index)."""
    if tail != "default":
        monkeypatch.setenv("TRT_TAIL_PATHS", tail)
    dev, orc = device_scenes[name], oracle_scenes[name]
    rays = make_rays(host_scenes[name], orc, 6000, seed=99)
    ids, t = dev.trace_closest(rays)
    got = dev.shade(rays, ids, t, seed=5, sample=3)
    ref = orc.shade_batch(rays, ids, t, seed=5, sample=3)
    assert got.shape == ref.shape and np.all(got[ids < 0] == 0)
    mean = ref.mean()
    assert mean > 0
    rmse = np.sqrt(((got.astype(np.float64) - ref) ** 2).mean())
    close = np.isclose(got, ref, rtol=1e-4, atol=1e-6).mean()
    assert rmse <= 1e-3 * mean and close >= 0.999, (name, rmse / mean, close)
    # truncated at depth 1 = direct light + emission only; another sample index gives other numbers
    d1 = dev.shade(rays, ids, t, seed=5, sample=3, max_depth=1)
    r1 = orc.shade_batch(rays, ids, t, seed=5, sample=3, max_depth=1)
    assert np.isclose(d1, r1, rtol=1e-4, atol=1e-6).mean() >= 0.999
    assert not np.array_equal(dev.shade(rays, ids, t, seed=5, sample=4), got)


def test_shade_rejects_foreign_triangle_ids(device_scenes):
    import tinyraytracing_b200 as trt

    dev = device_scenes["back"]
    rays = np.zeros((2, 6), np.float32)
    with pytest.raises(trt.TrtError):
        dev.shade(rays, np.array([0, 10 ** 6], np.int32), np.ones(2, np.float32))


def test_checkpoint_resume_equals_uninterrupted(tmp_path, device_scenes):
    """trt_accum_save after samples [0, 3), trt_accum_load into a fresh buffer, samples [3, 8): bit-identical to the same
    two accumulate calls without the file in between, and equal to trt_render up to the order of the double sums."""
    import tinyraytracing_b200 as trt

    dev = device_scenes["veach-mis"]
    spp, seed, path = 8, 13, str(tmp_path / "frame.ckpt")
    a = dev.accum_create()
    dev.render_accumulate(dev.params(spp, 0, 3, seed=seed), a)
    dev.accum_save(a, 3, spp, path, seed=seed)
    dev.render_accumulate(dev.params(spp, 3, spp, seed=seed), a)
    straight = dev.resolve(a, spp)
    dev.accum_destroy(a)

    b = dev.accum_create()
    info = dev.accum_load(path, b)
    assert info == dict(samples_done=3, spp=spp, seed=seed, max_depth=0)
    dev.render_accumulate(dev.params(spp, info["samples_done"], spp, seed=seed), b)
    resumed = dev.resolve(b, spp)
    assert np.array_equal(resumed, straight)
    assert np.allclose(resumed, dev.render(spp, seed=seed), rtol=1e-12, atol=0)

    # damaged / foreign files are errors, never a partial buffer
    raw = open(path, "rb").read()
    for bad in (raw[:-9], raw[:100] + bytes([raw[100] ^ 1]) + raw[101:], b"nonsense", raw + b"x"):
        with open(path, "wb") as f:
            f.write(bad)
        with pytest.raises(trt.TrtError):
            dev.accum_load(path, b)
    with open(path, "wb") as f:
        f.write(raw)
    other = device_scenes["back"]  # another frame size
    c = other.accum_create()
    with pytest.raises(trt.TrtError):
        other.accum_load(path, c)
    other.accum_destroy(c)
    dev.accum_destroy(b)
    assert not os.path.exists(path + ".tmp")


def test_render_multi_on_one_gpu(host_scenes, device_scenes):
    """The fan-out path with two replicas of the scene on the same GPU and the library's own peer reduce (NCCL refuses
    duplicate devices): sample ranges 4 + 3, summed in rank order and resolved by one kernel — the frame of trt_render."""
    import tinyraytracing_b200 as trt

    single = device_scenes["veach-mis"].render(7, seed=21)
    a = trt.DeviceScene(host_scenes["veach-mis"], 0)
    b = a.replicate(0)  # trt_scene_replicate: device-to-device copy of the uploaded scene, layouts not rebuilt
    try:
        probe = np.random.default_rng(0).normal(size=(2000, 6)).astype(np.float32)
        probe[:, :3] = probe[:, :3] * 0.5 + np.array(host_scenes["veach-mis"].camera()["eye"])
        ia, ta = a.trace_closest(probe)
        ib, tb = b.trace_closest(probe)
        assert np.array_equal(ia, ib) and np.array_equal(ta.view(np.uint32), tb.view(np.uint32)) and (ia >= 0).any()
        img, rgb = trt.render_multi([a, b], 7, seed=21, flags=trt.RENDER_PEER_REDUCE, want_rgb8=True)
        assert np.allclose(img, single, rtol=1e-12, atol=0)
        # main.cpp:30-38 (device pow vs glibc pow may truncate differently on an exact integer boundary)
        g = np.clip(np.power(img, np.float64(np.float32(1.0) / np.float32(2.2))) * 255, 0, 255).astype(np.uint8)
        assert np.abs(rgb.astype(int) - g.astype(int)).max() <= 1 and (rgb == g).mean() > 0.999
        assert a.stats()["paths"] == a.width * a.height * 4 and b.stats()["paths"] == a.width * a.height * 3
        # n = 1 is trt_render
        assert np.array_equal(trt.render_multi([a], 7, seed=21), single)
        with pytest.raises(trt.TrtError):  # NCCL path: duplicate devices are refused with a message, not a crash
            trt.render_multi([a, b], 7, seed=21)
        with pytest.raises(trt.TrtError):
            trt.render_multi([a, a], 7, seed=21, flags=trt.RENDER_PEER_REDUCE)
    finally:
        a.close()
        b.close()


@pytest.mark.parametrize("flags", (0, 8))
def test_render_multi_two_gpus(flags, host_scenes, device_scenes):
    """One process, two GPUs: ncclCommInitAll + ncclReduce(sum, f64) (flags 0) or the peer-memory kernel (flags 8)."""
    import torch

    import tinyraytracing_b200 as trt

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    single = device_scenes["veach-mis"].render(7, seed=21)
    devs = [trt.DeviceScene(host_scenes["veach-mis"], 0)]
    devs.append(devs[0].replicate(1))
    try:
        img = trt.render_multi(devs, 7, seed=21, flags=flags)
        assert np.allclose(img, single, rtol=1e-12, atol=0)
    finally:
        for d in devs:
            d.close()


def test_strict_walk_counter(host_scenes, device_scenes):
    """Rays from beyond 8 x scene scale (and non-finite ones) take the reference's exhaustive walk; trt_stats says how many."""
    dev, host = device_scenes["veach-mis"], host_scenes["veach-mis"]
    lo, hi = host.root_box()
    scale = float(np.abs(np.concatenate([lo, hi])).max())
    c = 0.5 * (lo + hi)
    rng = np.random.default_rng(1)
    o = c + 20 * scale * np.array([0.3, 0.4, -0.86])
    tgt = c + rng.uniform(-0.2, 0.2, (100, 3)) * (hi - lo)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    far = np.concatenate([np.tile(o, (100, 1)), d], 1).astype(np.float32)
    near = far.copy()
    near[:, :3] = (c + 0.5 * scale * np.array([0.3, 0.4, -0.86])).astype(np.float32)
    dev.reset_stats()
    dev.trace_closest(near)
    assert dev.stats()["rays_strict"] == 0
    ids, t = dev.trace_closest(far)
    assert dev.stats()["rays_strict"] == 100
    ex_ids, ex_t = dev.trace_closest(far, 2)  # TRT_TRACE_EXHAUSTIVE
    assert np.array_equal(ids, ex_ids) and np.array_equal(t.view(np.uint32), ex_t.view(np.uint32))
    dev.reset_stats()
    assert dev.stats()["rays_strict"] == 0


def test_render_multi_edge_cases(host_scenes, device_scenes):
    """Fewer samples than GPUs (one replica gets an empty range), a sample sub-range, rgb8 only, and a replica of a
    replica: the fan-out must still give trt_render's frame."""
    import ctypes as C

    import tinyraytracing_b200 as trt

    dev = device_scenes["back"]
    a = trt.DeviceScene(host_scenes["back"], 0)
    b = a.replicate(0)
    c = b.replicate(0)
    try:
        one = dev.render(1, seed=4)
        assert np.array_equal(trt.render_multi([a, b, c], 1, seed=4, flags=trt.RENDER_PEER_REDUCE), one)
        assert b.stats()["paths"] == 0 or a.stats()["paths"] == 0 or c.stats()["paths"] == 0  # somebody had nothing to do
        five = dev.render(5, seed=4)
        got = trt.render_multi([a, b, c], 5, seed=4, flags=trt.RENDER_PEER_REDUCE)
        assert np.allclose(got, five, rtol=1e-12, atol=0)
        # image_rgb = NULL, rgb8 only
        rgb = np.empty((a.height, a.width, 3), np.uint8)
        handles = (C.c_void_p * 2)(a.h, b.h)
        p = a.params(5, 0, 5, seed=4, flags=trt.RENDER_PEER_REDUCE)
        assert a.lib.trt_render_multi(handles, 2, C.byref(p), None, rgb.ctypes.data) == 0
        ref = np.clip(np.power(five, np.float64(np.float32(1.0) / np.float32(2.2))) * 255, 0, 255).astype(np.uint8)
        assert np.abs(rgb.astype(int) - ref.astype(int)).max() <= 1
        # bad sample range
        p = a.params(5, 3, 2, seed=4)
        assert a.lib.trt_render_multi(handles, 2, C.byref(p), None, rgb.ctypes.data) == -1
    finally:
        for d in (a, b, c):
            d.close()


def test_shade_of_nothing_and_of_misses(device_scenes):
    dev = device_scenes["veach-mis"]
    assert dev.shade(np.zeros((0, 6), np.float32), np.zeros(0, np.int32), np.zeros(0, np.float32)).shape == (0, 3)
    rays = np.tile(np.array([0, 0, 0, 0, 1, 0], np.float32), (5, 1))
    out = dev.shade(rays, np.full(5, -1, np.int32), np.full(5, 114514.0, np.float32))
    assert np.all(out == 0)


def test_trace_closest_multi_shards_by_ray_index(host_scenes, oracle_scenes, device_scenes):
    """One host batch spread over replicas (here three on one GPU; ragged shard sizes, a shard of zero rays): ids and
    distance bits of the single-scene call, which are the oracle's."""
    import tinyraytracing_b200 as trt

    a = trt.DeviceScene(host_scenes["staircase"], 0)
    reps = [a, a.replicate(0), a.replicate(0)]
    try:
        rays = make_rays(host_scenes["staircase"], oracle_scenes["staircase"], 100003, seed=12)
        ids, t = trt.trace_closest_multi(reps, rays)
        oid, ot = oracle_scenes["staircase"].trace(rays)
        assert np.array_equal(ids, oid) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
        ids2, t2 = trt.trace_closest_multi(reps, rays[:2])  # fewer rays than scenes
        assert np.array_equal(ids2, oid[:2]) and np.array_equal(t2.view(np.uint32), ot[:2].view(np.uint32))
        with pytest.raises(trt.TrtError):
            trt.trace_closest_multi([a, a], rays[:10])
    finally:
        for d in reps:
            d.close()


def test_scene_created_from_the_layout_cache_is_the_same_scene(host_scenes, device_scenes, tmp_path):
    """trt_scene_create_cached: the first call builds the layouts and writes the file, the second reads it; ids, distance
    bits and the frame are those of the scene that trt_scene_create builds."""
    import tinyraytracing_b200 as trt

    host, plain = host_scenes["staircase"], device_scenes["staircase"]
    path = str(tmp_path / "staircase.layout")
    rng = np.random.default_rng(5)
    lo, hi = host.root_box()
    rays = np.concatenate([rng.uniform(lo, hi, (40000, 3)), rng.normal(size=(40000, 3))], 1).astype(np.float32)
    ref_ids, ref_t = plain.trace_closest(rays)
    ref_img = plain.render(3, seed=8)
    assert (ref_ids >= 0).mean() > 0.3
    for expect_hit in (False, True):
        dev = trt.DeviceScene(host, 0, layout_cache=path)
        try:
            assert dev.layout_from_cache == expect_hit
            ids, t = dev.trace_closest(rays)
            assert np.array_equal(ids, ref_ids) and np.array_equal(t.view(np.uint32), ref_t.view(np.uint32))
            assert np.array_equal(dev.render(3, seed=8), ref_img)
            assert dev.stats()["accel_nodes"] == plain.stats()["accel_nodes"]
        finally:
            dev.close()
