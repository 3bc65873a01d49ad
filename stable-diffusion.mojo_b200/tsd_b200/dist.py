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


class DistC:
    """The C ABI's own multi-GPU entry points (tsd_dist_*, csrc/dist_nccl.cu): one NCCL communicator per process, bound at
    run time, no torch.distributed involved.  `id_bytes` is the 128-byte id of `DistC.unique_id()` (rank 0) handed to the
    other ranks by the host program, or None with a `rendezvous` file path every rank can reach."""

    def __init__(self, ctx, nranks: int, rank: int, id_bytes: bytes | None = None, rendezvous: str | None = None):
        import ctypes as C
        self.ctx, self.nranks, self.rank = ctx, nranks, rank
        d = C.c_void_p()
        idp = None if id_bytes is None else C.c_char_p(bytes(id_bytes))
        ctx._ck(ctx.L.tsd_dist_init(ctx.h, nranks, rank, idp, None if rendezvous is None else rendezvous.encode(), C.byref(d)))
        self.d = d

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from . import _lib
        buf = C.create_string_buffer(128)
        rc = _lib.lib().tsd_dist_unique_id(buf)
        if rc != 0:
            raise _lib.TsdError(rc, "tsd_dist_unique_id failed (NCCL not loadable?)")
        return buf.raw

    def broadcast_context(self, context: np.ndarray | None, shape, root: int = 0) -> np.ndarray:
        buf = np.ascontiguousarray(context, np.float32) if self.rank == root else np.empty(tuple(shape), np.float32)
        if tuple(buf.shape) != tuple(shape):
            raise ValueError("context shape mismatch")
        self.ctx._ck(self.ctx.L.tsd_dist_broadcast_context(self.d, buf.ctypes.data, buf.size, root))
        return buf

    def gather(self, local: np.ndarray, root: int = 0) -> np.ndarray | None:
        local = np.ascontiguousarray(local, np.float32)
        out = np.empty((self.nranks,) + local.shape, np.float32) if self.rank == root else None
        self.ctx._ck(self.ctx.L.tsd_dist_gather(self.d, local.ctypes.data, local.size,
                                                None if out is None else out.ctypes.data, root))
        return out

    def generate(self, diffusion, latents, context, timesteps, time_emb, coef, noise=None, cfg=False, cfg_scale=7.5,
                 n_ctx: int = 1, root: int = 0):
        """tsd_dist_generate: the context rows (valid on `root`) are broadcast, then this rank's latents are denoised."""
        import ctypes as C
        from . import _lib
        latents = np.ascontiguousarray(latents, np.float32)
        row = (77, 768)
        cbuf = np.ascontiguousarray(context, np.float32) if self.rank == root else np.empty((n_ctx,) + row, np.float32)
        timesteps = np.ascontiguousarray(timesteps, np.int32)
        time_emb = np.ascontiguousarray(time_emb, np.float32)
        coef = np.ascontiguousarray(coef, np.float32)
        nz = None if noise is None else np.ascontiguousarray(noise, np.float32)
        lp = _lib.LoopParams(len(timesteps), int(cfg), float(cfg_scale), timesteps.ctypes.data_as(_lib.c_i32_p),
                             time_emb.ctypes.data_as(_lib.c_float_p), coef.ctypes.data_as(_lib.c_float_p),
                             None if nz is None else nz.ctypes.data_as(_lib.c_float_p))
        out = np.empty_like(latents)
        self.ctx._ck(self.ctx.L.tsd_dist_generate(self.d, diffusion.m, C.byref(lp), latents.ctypes.data, cbuf.ctypes.data,
                                                  cbuf.shape[0], latents.shape[0], root, out.ctypes.data))
        return out, cbuf

    def close(self):
        if getattr(self, "d", None):
            self.ctx.L.tsd_dist_shutdown(self.d)
            self.d = None
