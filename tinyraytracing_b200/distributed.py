"""Multi-GPU plumbing (one process per GPU, torch.distributed): sample-range sharding with the scene replicated
and ONE sum-reduce of the per-pixel accumulation buffers.  The path has no other exchange step:
samples are independent given the counter-based RNG keyed on (pixel, sample, depth, slot).

torch is plumbing here (device memory for the accumulation buffer, the NCCL reduce over NVLink); the
rendering itself happens inside libtrt_b200.so."""
import numpy as np


def shard_samples(spp, world, rank):
    """Contiguous sample range [lo, hi) of `rank`: sizes differ by at most one, ranges tile [0, spp)."""
    base, rem = divmod(spp, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def render_distributed(accumulate, resolve, shape, spp, device=None, group=None, dst=0):
    """accumulate(lo, hi, acc_tensor) adds the radiance sums of samples [lo, hi) into acc (float64, H*W*3);
    resolve(acc_tensor) -> image on the destination rank.  Returns the image on `dst`, None elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    acc = torch.zeros(int(np.prod(shape)), dtype=torch.float64, device=device)
    lo, hi = shard_samples(spp, world, rank)
    if hi > lo:
        accumulate(lo, hi, acc)
    if world > 1:
        dist.reduce(acc, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return resolve(acc) if rank == dst else None


def render_on_gpus(dev, spp, seed=0, max_depth=0, group=None, out=None):
    """The reference's image (float64 H x W x 3, divided by spp) on rank 0, rendered by all ranks' GPUs.
    out: optional destination array (e.g. dev.pinned_image()), see DeviceScene.resolve."""
    import torch

    stream = torch.cuda.current_stream().cuda_stream

    def accumulate(lo, hi, acc):
        dev.render_accumulate(dev.params(spp, lo, hi, max_depth=max_depth, seed=seed), acc.data_ptr(), stream)

    def resolve(acc):
        return dev.resolve(acc.data_ptr(), spp, stream=stream, out=out)

    return render_distributed(accumulate, resolve, (dev.height, dev.width, 3), spp,
                              device=torch.device("cuda", torch.cuda.current_device()), group=group)
