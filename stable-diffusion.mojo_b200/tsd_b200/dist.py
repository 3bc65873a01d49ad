"""Multi-GPU plumbing for the denoising loop: batch (data) sharding, one process per GPU.

The path shards by independent images (the reference's only batching hint is the
comment at pipeline.mojo:12).  Rank r owns samples r, r+R, r+2R, ...; every rank holds a full
weight replica; seeds derive from the SAMPLE index so results do not depend on the rank count.
The only collectives are one broadcast of the CLIP context per prompt and an optional gather of
the decoded images - there is NO per-step collective.  torch.distributed is plumbing only
(NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import os

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_indices(global_batch: int, rank: int, world: int) -> list[int]:
    """Samples owned by `rank`: r, r+R, ... (round-robin keeps ranks within one sample of each other)."""
    if global_batch < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad shard arguments")
    return list(range(rank, global_batch, world))


def sample_seed(base_seed: int, sample_index: int) -> int:
    """Seed of one image: a function of the sample index only (rank-count invariant)."""
    return int(base_seed) + 1000 * int(sample_index)


def sample_inputs(base_seed: int, sample_index: int, side: int, steps: int):
    """Initial latent (4,side,side) and per-step noise (steps,4,side,side), N(0,1)."""
    rng = np.random.default_rng(sample_seed(base_seed, sample_index))
    lat = rng.standard_normal((4, side, side), dtype=np.float32)
    noise = rng.standard_normal((steps, 4, side, side), dtype=np.float32)
    return lat, noise


def init_process_group(backend: str | None = None):
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def _device_for_backend():
    import torch
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_context(context: np.ndarray | None, shape, src: int = 0) -> np.ndarray:
    """One broadcast per prompt of the (n_ctx,77,768) context (236 544 B per row) from `src`."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.ascontiguousarray(context, np.float32)
    dev = _device_for_backend()
    if dist.get_rank() == src:
        t = torch.from_numpy(np.ascontiguousarray(context, np.float32)).to(dev)
        if tuple(t.shape) != tuple(shape):
            raise ValueError("context shape mismatch")
    else:
        t = torch.empty(tuple(shape), dtype=torch.float32, device=dev)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def gather_samples(local: np.ndarray, global_batch: int) -> np.ndarray | None:
    """Gathers per-rank results [(n_local, ...)] into sample order on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device_for_backend()
    per = (global_batch + world - 1) // world
    buf = np.zeros((per,) + local.shape[1:], np.float32)
    buf[:local.shape[0]] = local
    t = torch.from_numpy(buf).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    if rank != 0:
        return None
    full = np.empty((global_batch,) + local.shape[1:], np.float32)
    for r in range(world):
        idx = shard_indices(global_batch, r, world)
        full[idx] = outs[r].cpu().numpy()[:len(idx)]
    return full


def max_over_ranks(value: float) -> float:
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
