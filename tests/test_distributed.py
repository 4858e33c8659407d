"""N > 1 path on CPU: world_size-2 gloo run of the sample-range sharding + sum-reduce (the plumbing of
tinyraytracing_b200.distributed), with the oracle standing in for the per-rank renderer."""
import os
import socket

import numpy as np
import pytest

from tinyraytracing_b200.distributed import shard_samples


def test_shard_samples_tile_the_range():
    for spp in (1, 2, 7, 16, 1024):
        for world in (1, 2, 3, 4, 8):
            r = [shard_samples(spp, world, k) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == spp
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, spp, out):
    import torch
    import torch.distributed as dist

    import oraclelib
    from tinyraytracing_b200.distributed import render_distributed

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = oraclelib.OracleScene(oraclelib.parsed_scene("back", 24, 24))

    def accumulate(lo, hi, acc):
        img, _ = o.render(spp, seed=5, sample_begin=lo, sample_end=hi, threads=1)
        acc += torch.from_numpy(img.reshape(-1) * spp)  # radiance sums, like trt_render_accumulate

    img = render_distributed(accumulate, lambda acc: (acc / spp).numpy().reshape(24, 24, 3), (24, 24, 3), spp)
    if rank == 0:
        np.save(out, img)
    else:
        assert img is None
    dist.destroy_process_group()


def test_two_rank_gloo_render_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp

    import oraclelib

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "img.npy")
    spp = 5
    mp.spawn(_worker, args=(2, port, spp, out), nprocs=2, join=True)
    ref, _ = oraclelib.OracleScene(oraclelib.parsed_scene("back", 24, 24)).render(spp, seed=5)
    assert np.allclose(np.load(out), ref, rtol=1e-12, atol=1e-15)
