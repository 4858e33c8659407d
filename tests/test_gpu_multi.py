"""Multi-GPU path on real devices: one process per GPU, scene replicated, sample ranges sharded, one NCCL
sum-reduce of the accumulation buffers.  Skipped on a box with a single GPU (the CPU suite covers the sharding
logic with gloo, tests/test_distributed.py)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, spp, out, files):
    import torch
    import torch.distributed as dist

    import tinyraytracing_b200 as trt
    from tinyraytracing_b200.distributed import render_on_gpus

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    host = trt.HostScene.load(files["xml"], files["obj"], files["mtl"], files["basedir"])
    dev = trt.DeviceScene(host, rank)
    img = render_on_gpus(dev, spp, seed=11)
    if rank == 0:
        np.save(out, img)
    else:
        assert img is None
    dev.close()
    dist.destroy_process_group()


def test_two_gpu_render_equals_single_gpu(tmp_path, scene_files, device_scenes):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "img.npy")
    spp = 7  # odd: uneven shards (4 + 3)
    mp.spawn(_worker, args=(2, port, spp, out, scene_files["veach-mis"]), nprocs=2, join=True)
    single = device_scenes["veach-mis"].render(spp, seed=11)
    assert np.allclose(np.load(out), single, rtol=1e-12, atol=0)
