"""Multi-GPU check of the C ABI's own entry points (tsd_dist_*, csrc/dist_nccl.cu): run one process per GPU, e.g.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_c_abi_check.py

torchrun is only the launcher here (RANK / WORLD_SIZE / LOCAL_RANK): the communicator is NCCL bound by the library at run
time, bootstrapped through a rendezvous file; torch.distributed is never initialised.  Every rank denoises its own latent
(batch sharding, no per-step collective) with the context broadcast from rank 0; rank 0 gathers the latents and checks
them against the same samples evaluated in one process."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion.mojo_b200"))
from tsd_b200 import dist as tdist  # noqa: E402
from tsd_b200.api import Context, Diffusion  # noqa: E402
from tsd_b200.sampler import DDPMSampler, get_time_embedding  # noqa: E402


def main():
    rank, world, local = tdist.env_rank_world()
    rdv = os.environ.get("TSD_DIST_RENDEZVOUS", f"/tmp/tsd_rdv_{os.environ.get('MASTER_PORT', '0')}")
    if rank == 0 and os.path.exists(rdv):
        os.remove(rdv)
    ctx = Context(local)
    d = tdist.DistC(ctx, world, rank, rendezvous=rdv)
    side, steps = 8, 3
    m = Diffusion(ctx, side, side, max_batch=1)
    m.init_random(1234)
    sm = DDPMSampler()
    sm.set_inference_timesteps(steps)
    ts = sm.timesteps.astype(np.int32)
    temb = np.stack([get_time_embedding(float(t)) for t in ts]).astype(np.float32)
    coef = sm.coefficient_table()
    context = np.random.default_rng(99).standard_normal((1, 77, 768), dtype=np.float32) if rank == 0 else None
    lat0, noise = tdist.sample_inputs(1234, rank, side, steps)
    # tsd_dist_generate: broadcast + this rank's loop
    lat, ctx_rows = d.generate(m, lat0[None], context, ts, temb, coef, noise[:, None], n_ctx=1, root=0)
    want_ctx = np.random.default_rng(99).standard_normal((1, 77, 768), dtype=np.float32)
    assert np.array_equal(ctx_rows, want_ctx), "context broadcast mismatch"
    assert np.array_equal(d.broadcast_context(context if rank == 0 else None, (1, 77, 768)), want_ctx)
    allv = d.gather(lat[0], root=0)
    ok = True
    if rank == 0:
        for r in range(world):
            l0, nz = tdist.sample_inputs(1234, r, side, steps)
            single = m.generate_latents(l0[None], want_ctx, ts, temb, coef, nz[:, None])
            err = float(np.abs(allv[r] - single[0]).max() / np.abs(single[0]).max())
            ok = ok and err < 5e-3          # same kernels, another process / GPU: TF32-level agreement
            print(f"rank {r}: gathered latent vs single-process evaluation rel_linf {err:.2e}")
        print("DIST_C_ABI_OK" if ok else "DIST_C_ABI_FAIL", flush=True)
    d.close()
    m.close()
    ctx.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
