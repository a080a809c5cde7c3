"""Multi-GPU plumbing for the one place the hot path shards: independent (latent, camera, mesh) samples.

One process per GPU (torchrun), contiguous split of the sample list across ranks, no data-path collective inside the
generator; a single gather of the output images to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests) that can be
issued on a side stream so it overlaps the next micro-batch.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank). No-op for world 1."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        # NCCL writes its debug output (the version banner included, at any NCCL_DEBUG level) to STDOUT unless told otherwise:
        # send it to stderr so that a caller's stdout stays machine-readable (bench.py prints one JSON line)
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(total, rank, world):
    """Contiguous [start, stop) of `total` samples owned by `rank` (first `total % world` ranks get one extra)."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_images(local_images, dst=0, sizes=None):
    """Gather per-rank image batches to `dst`: returns [sum(B_r), ...] on dst (rank order), None elsewhere.  `sizes`: batch size of
    every rank when the shards are unequal (shard_range with total % world != 0) -- smaller shards are padded to the largest for
    the collective and trimmed on dst; None = all ranks hold the same number of images."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_images
    world, rank = dist.get_world_size(), dist.get_rank()
    bmax = local_images.shape[0] if sizes is None else max(sizes)
    send = local_images.contiguous()
    if send.shape[0] < bmax:
        send = torch.cat([send, send.new_zeros((bmax - send.shape[0],) + tuple(send.shape[1:]))], 0)
    if rank == dst:
        out = torch.empty((world,) + tuple(send.shape), dtype=send.dtype, device=send.device)
        dist.gather(send, list(out.unbind(0)), dst=dst)
        if sizes is None:
            return out.flatten(0, 1)
        return torch.cat([out[r, :sizes[r]] for r in range(world)], 0)
    dist.gather(send, None, dst=dst)
    return None


def max_over_ranks(value, device):
    """Scalar max over ranks (device timing of multi-GPU runs is the slowest rank's)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
