#!/usr/bin/env python
"""bench.py - the reference's headline workloads on B200 (BASELINE.json configs[1..4]).

    python bench.py --gpus N --steps K --warmup W                    # configs[1]: UNet denoising steps/s (headline)
    python bench.py --config cfg50 --gpus N [--steps K --warmup W]   # configs[2]: 50 steps, CFG 7.5, one image per GPU
    python bench.py --config vae16 [--steps K --warmup W]            # configs[3]: VAE decoder, batch 16
    python bench.py --config attn  [--steps K --warmup W]            # configs[4]: attention core sweep vs cuBLAS / SDPA
    python bench.py --impl reference [--config ...] ...              # CPU restatement of the reference, all host cores

configs[1] (default, `--config unet20`): one bench "step" = one iteration of the reference denoising loop without CFG
(pipeline.mojo:86-122): Diffusion.forward (diffusion.mojo:309-318) on a (4,64,64) latent with a (1,77,768) context and
a (320,) time embedding, followed by DDPMSampler.step (sampler.mojo:75-109).  With N GPUs every rank owns one latent
(batch sharding, no per-step collective) and the context is NCCL-broadcast once before the loop.

Keys of the JSON line (see DESIGN.md "Measurement"):
  value      whole-job throughput with the inputs resident in HBM (CUDA events on the library's stream)
  e2e        the same metric through the host-buffer C ABI with pinned host buffers, H2D/D2H inside the timed region
             (unet20: tsd_diffusion_step = Diffusion.forward + DDPMSampler.step in one call; the K/V projections of
             an unchanged context are reused)
  roofline   dominant kernel family: algorithmic FLOPs / CUDA-event time vs the TF32 peak (burst AND sustained
             denominators are printed; `frac` uses the burst one: the timed regions are milliseconds long)
  cpu_baseline  oracle/ref_loops.c (C restatement of the reference's scalar loops, OpenMP on ALL host cores)
Timing: barrier + synchronize on both sides, max over ranks.  L2: the 1.2 GB fp32 weight stream per UNet step is ~10x
the 126 MB L2 (no flush needed); the other configs say how they handle it in config.l2.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "stable-diffusion.mojo_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "unet_denoising_steps_per_sec_512x512_bs1"
UNIT = "steps/s"
UNET_GFLOP = 408.33          # one Diffusion.forward at a 4x64x64 latent, 77x768 context
CTX_KV_GFLOP = 1.59          # the 18 context K/V projections (M=77), hoisted out of the loop
VAE_GFLOP = 2514.52          # one Decoder.forward 4x64x64 -> 3x512x512
SIDE, CTX_LEN, CTX_DIM, LOOP_STEPS = 64, 77, 768, 20
WORKLOADS = {
    "unet20": "Tiny-SD 512x512 txt2img, 20 DDPM steps, batch 1, fp32, 1xB200 (BASELINE configs[1])",
    "cfg50": "Tiny-SD 512x512, 50 steps, CFG (uncond+cond), one image per GPU (BASELINE configs[2])",
    "vae16": "VAE decoder only, 64x64x4 -> 512x512x3, batch 16, 1xB200 (BASELINE configs[3])",
    "attn": "Attention kernel sweep: seq 4096, ctx 77, d_head 40/80/160 vs cuBLAS / SDPA (BASELINE configs[4])",
}
WORKLOAD = WORKLOADS["unet20"]
DTYPE = "tf32 (fp32 storage, tcgen05 kind::tf32 products, fp32 accumulate)"


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def base_config(n_gpus, name="unet20"):
    cfg = {"workload": WORKLOADS[name], "name": name}
    if name == "unet20":
        cfg.update({"latent": [4, SIDE, SIDE], "context": [1, CTX_LEN, CTX_DIM], "per_gpu_batch": 1,
                    "global_batch": n_gpus, "cfg": False,
                    "sharding": "one latent per rank, context broadcast once, no per-step collective",
                    "weights": "synthetic seeded (reference init ranges), 299.74 M params fp32",
                    "context_kv": "hoisted once per prompt in `value`; reused while the context bytes are unchanged in `e2e`",
                    "l2": "no flush: 1.2 GB weight stream per step >> 126 MB L2"})
    elif name == "cfg50":
        cfg.update({"latent": [4, SIDE, SIDE], "context": [2, CTX_LEN, CTX_DIM], "steps_per_image": 50, "cfg_scale": 7.5,
                    "per_gpu_batch": 1, "global_batch": n_gpus,
                    "sharding": "one image per rank (UNet batch 2 = cond + uncond), context broadcast once, no per-step collective",
                    "l2": "no flush: 1.2 GB weight stream per UNet evaluation >> 126 MB L2"})
    elif name == "vae16":
        cfg.update({"latent": [16, 4, SIDE, SIDE], "image": [16, 3, 8 * SIDE, 8 * SIDE], "per_gpu_batch": 16,
                    "global_batch": 16 * n_gpus,
                    "l2": "no flush: activations of up to 2.1 GB per layer at batch 16 >> 126 MB L2"})
    else:
        cfg.update({"heads": 8, "tq": 4096, "tk": [4096, 77], "d_head": [40, 80, 160],
                    "l2": "q/k/v of one case are 5-84 MB: L2-resident between iterations, as inside the UNet step"})
    return cfg


# ------------------------------------------------------------------------------------------------
# CPU arm: oracle/ref_loops.c (restatement of the reference's scalar loops, OpenMP over the axes the
# reference hands to `parallelize`) through oracle/tsd_oracle.py Ops("c32").
# ------------------------------------------------------------------------------------------------
def cpu_threads_all():
    """The CPU baseline is defined on ALL host cores.  torchrun exports OMP_NUM_THREADS=1 to every rank: set the
    OpenMP team size explicitly instead of inheriting it."""
    import tsd_oracle as O
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return int(O.clib().ref_set_num_threads(0))


class CpuSample:
    """Composite slice of one Diffusion.forward with the step's conv/linear/attention FLOP mix
    (47.8/33.7/18.5 %): at scale s = 1 a 3x3 conv 320->320 at 64x64 (UNet layer2
    shape), a 3248x320 -> 2560 Linear (GEGLU shape) and one T=4096, d=40 attention head.  Used for warm-up and as a
    cross-check of the full step; never the reported number when a full step fits the time budget."""

    def __init__(self):
        import tsd_oracle as O
        self.O = O
        self.cores = cpu_threads_all()
        self.ops = O.Ops("c32")
        rng = np.random.default_rng(7)
        self.x = rng.standard_normal((320, 64, 64), dtype=np.float32)
        self.w = (rng.standard_normal((320, 320, 3, 3)) / np.sqrt(2880)).astype(np.float32)
        self.b = np.zeros(320, np.float32)
        self.lx = rng.standard_normal((3248, 320), dtype=np.float32)
        self.lw = (rng.standard_normal((2560, 320)) / np.sqrt(320)).astype(np.float32)
        self.lb = rng.standard_normal(2560, dtype=np.float32)
        self.q = rng.standard_normal((1, 4096, 40), dtype=np.float32)
        self.k = rng.standard_normal((1, 4096, 40), dtype=np.float32)
        self.v = rng.standard_normal((1, 4096, 40), dtype=np.float32)

    def shape(self, s):
        oc = max(1, int(round(320 * s)))
        rows = max(1, int(round(3248 * s)))
        t = max(16, int(round(4096 * np.sqrt(s))))
        gflop = (2 * 4096 * oc * 2880 + 2 * rows * 320 * 2560 + 4 * t * t * 40) / 1e9
        return oc, rows, min(t, 4096), gflop

    def run(self, s):
        """Runs the sample once; returns (seconds, GFLOP)."""
        oc, rows, t, gflop = self.shape(s)
        t0 = time.perf_counter()
        self.ops.conv2d(self.x, self.w[:oc], self.b[:oc], pad=1)
        self.ops.linear(self.lx[:rows], self.lw, self.lb)
        self.ops.attention_core(self.q[:, :t], self.k[:, :t], self.v[:, :t])
        return time.perf_counter() - t0, gflop

    def describe(self, s, reps):
        oc, rows, t, gflop = self.shape(s)
        return (f"{reps}x [conv3x3 320->{oc} @64x64 + linear {rows}x320->2560 + attention core h=1 T={t} d=40] "
                f"= {gflop:.2f} GFLOP of the {UNET_GFLOP} GFLOP step (same conv/linear/attention mix), "
                f"scaled by FLOPs to one full UNet step; oracle/ref_loops.c, OpenMP")

    def estimate_step_seconds(self, budget_s=4.0):
        """FLOP-scaled estimate of one full UNet step from a ~budget_s composite sample (also warms the thread pool)."""
        sec, g = self.run(0.01)
        rate = g / max(sec, 1e-6)
        s = float(min(1.0, max(0.004, budget_s * rate / self.shape(1.0)[3])))
        sec, g = self.run(s)
        return sec / g * UNET_GFLOP, s


class CpuFullStep:
    """One REAL UNet step of the restated reference on the CPU: oracle diffusion_forward(Ops("c32")) - every layer of
    diffusion.mojo:228-291 through the C loop nests (conv, matmul, column softmax, GroupNorm, ...) - at the 64x64
    latent of BASELINE configs[1], followed by the sampler step."""

    def __init__(self):
        import synth
        import tsd_oracle as O
        self.O = O
        self.ops = O.Ops("c32")
        self.W = synth.SynthWeights(synth.diffusion_specs(), 1234)
        rng = np.random.default_rng(5)
        self.lat = rng.standard_normal((4, SIDE, SIDE), dtype=np.float32)
        self.ctx = rng.standard_normal((CTX_LEN, CTX_DIM), dtype=np.float32)
        self.noise = rng.standard_normal((4, SIDE, SIDE), dtype=np.float32)
        self.sm = O.DDPMSampler()
        self.sm.set_inference_timesteps(LOOP_STEPS)

    def run(self, i=0):
        t = int(self.sm.timesteps[i % LOOP_STEPS])
        t0 = time.perf_counter()
        eps = self.O.diffusion_forward(self.ops, self.W, self.lat, self.ctx, self.O.get_time_embedding(float(t)))
        self.sm.step(t, self.lat, eps, self.noise)
        return time.perf_counter() - t0


def cpu_full_steps(budget_s, max_steps, min_steps=1):
    """Times real UNet steps on all host cores until `max_steps` are done or the budget is used (at least `min_steps`).
    Returns (steps/s, steps timed, seconds, cores, composite cross-check)."""
    sample = CpuSample()
    est, s = sample.estimate_step_seconds()
    full = CpuFullStep()
    secs = [full.run(0)]           # the first full step also sizes the rest of the run
    n = int(max(min_steps, min(max_steps, budget_s // max(secs[0], 1e-3))))
    secs += [full.run(i) for i in range(1, n)]
    total = float(sum(secs))
    return n / total, n, total, sample.cores, {"composite_estimate_steps_per_s": 1.0 / est, "composite_scale": s,
                                               "composite_sample": sample.describe(s, 1)}


def run_reference(args):
    """The reference arm: the oracle's C restatement on all host cores (the Mojo 24.x reference cannot be built here).
    Every timed step is one REAL Diffusion.forward + sampler step at the full 64x64 latent; because one such step takes
    tens of seconds, the arm times as many as fit its budget (at least one) and reports that count in `steps`."""
    rank, world, _ = env_rank()
    if rank != 0:
        return
    name = args.config
    budget = float(os.environ.get("TSD_REF_BUDGET_S", "150"))
    t_wall = time.perf_counter()
    value, n, secs, cores, cross = cpu_full_steps(budget, max(1, args.steps))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n, "steps_requested": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference", "extrapolated": False,
            "config": base_config(args.gpus, name)}
    sample_txt = (f"{n} full UNet step(s): oracle diffusion_forward(Ops('c32')) = every layer of diffusion.mojo:228-291 at the "
                  f"4x64x64 latent + DDPMSampler.step, oracle/ref_loops.c on {cores} OpenMP threads ({secs:.1f} s); warm-up = "
                  f"a bounded composite sample (conv/linear/attention mix)")
    if name == "cfg50":
        # one image = 100 UNet evaluations + 1 decode; the decode (2514.52 GFLOP, 98 % conv) is scaled from the UNet's
        # measured GFLOP/s - labelled as an extrapolation
        unet_s = 1.0 / value
        img_s = 100 * unet_s + VAE_GFLOP / (UNET_GFLOP * value)
        line.update(metric="images_per_sec_512x512_50steps_cfg", unit="images/s", value=1.0 / img_s,
                    ms_per_step=1000.0 * img_s, extrapolated=True)
        sample_txt += "; one image = 100 such evaluations + one VAE decode scaled by FLOPs (extrapolated: true)"
    elif name == "vae16":
        img_s = VAE_GFLOP / (UNET_GFLOP * value)
        line.update(metric="vae_decode_images_per_sec_512x512_bs16", unit="images/s", value=1.0 / img_s,
                    ms_per_step=16 * 1000.0 * img_s, extrapolated=True)
        sample_txt += "; decoder images/s scaled by FLOPs from the UNet's measured GFLOP/s (extrapolated: true)"
    elif name == "attn":
        import tsd_oracle as O
        ops = O.Ops("c32")
        rng = np.random.default_rng(3)
        q, k, v = (rng.standard_normal((8, 1024, 40), dtype=np.float32) for _ in range(3))
        t0 = time.perf_counter()
        ops.attention_core(q, k, v)
        dt = time.perf_counter() - t0
        tf = 4 * 8 * 1024 * 1024 * 40 / dt / 1e12
        line.update(metric="attention_core_tflops_T4096_d40", unit="TFLOP/s", value=tf, ms_per_step=dt * 1e3 * 16,
                    extrapolated=True)
        sample_txt = (f"attention core h=8 T=1024 d=40 (1/16 of the T=4096 case) in {dt:.2f} s on {cores} threads, "
                      f"oracle/ref_loops.c; TFLOP/s carried over to T=4096 (extrapolated: true)")
    line["cpu_baseline"] = {"value": line["value"], "unit": line["unit"], "cores": cores, "kind": "port",
                            "sample": sample_txt, "sample_seconds": secs, **cross}
    line["e2e"] = {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    line["gpu_launches"] = 0
    line["wall_s"] = time.perf_counter() - t_wall
    line["note"] = ("reference = the oracle's C restatement of the reference's scalar loops (the Mojo 24.x reference cannot be "
                    "built: no Mojo toolchain); `steps` = full UNet steps actually timed inside the time budget "
                    "(TSD_REF_BUDGET_S, default 150 s); at N > 1 rank 0 alone runs: the host's cores do not multiply with the "
                    "GPU count, so the CPU's whole-job steps/s is the same number")
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.path = f"/tmp/tsd_clocks_{os.getpid()}.csv"
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in open(self.path):
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def tf32_peaks():
    """TF32 dense peaks = 1/2 of the measured cuBLAS bf16 figures in MEASURED_PEAKS.json (tcgen05 kind::tf32 runs at half
    the kind::f16 rate, nominal 1.1 vs 2.25 PFLOP/s; the file has no TF32 row).  Returns (burst, sustained, source)."""
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return (0.5 * float(pk["bf16_tflops"]), 0.5 * float(pk["bf16_tflops_sustained"]),
                "0.5 x bf16_tflops (burst) / bf16_tflops_sustained of measured (MEASURED_PEAKS.json)")
    except Exception:
        return 0.5 * 1590.0, 0.5 * 1400.0, "0.5 x 1.59 / 1.4 PFLOP/s bf16 of fallback (B200_PROFILING.md)"


def own_tf32_measurement():
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_measured_tf32_peaks.json")))
        return t
    except Exception:
        return None


def kernel_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed
    `ncu --set full` capture (profiles/*_kernel_traffic.json, written by tools/ncu_summary.py from
    the capture named there); None when the file is absent."""
    for name in ("r02_kernel_traffic.json", "r01_kernel_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))
            return float(t["dram_bytes_per_launch"]), t.get("note")
        except Exception:
            continue
    return None, None


def setup_gpu(args):
    import torch
    from tsd_b200 import dist as tdist
    from tsd_b200.api import Context
    rank, world, local = env_rank()
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - this repo has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    tdist.init_process_group("nccl" if world > 1 else None)
    ctx = Context(local)          # raises if libtsd_b200.so or an sm_100 device is missing
    return torch, tdist, ctx, rank, world, local


def finish(ctx):
    import torch.distributed as dist
    ctx.close()
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def graph_timeline(run_once, sync, gemm_flops):
    """Exposed time per kernel family inside REPLAYS OF THE CAPTURED GRAPH (CUPTI activity records through torch.profiler):
    what a kernel adds to the critical path after its stream predecessor has ended - programmatic dependent launch lets a
    kernel start (and wait) while the previous one is still running, which the eager per-launch events above cannot see.
    Explanatory only: the profiler adds a little overhead, so `frac` in the roofline block stays the event-based figure."""
    try:
        import re
        import torch
        from torch.profiler import profile, ProfilerActivity
        for _ in range(3):
            run_once()
        sync()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(4):
                run_once()
                sync()
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA
              and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
        ev.sort(key=lambda e: e.time_range.start)
        runs, cur = [], []
        for e in ev:
            if cur and e.time_range.start - max(x.time_range.end for x in cur[-4:]) > 30:
                runs.append(cur)
                cur = []
            cur.append(e)
        runs.append(cur)
        runs = [r for r in runs if len(r) > 50]
        if not runs:
            return {"error": "no kernel records"}
        run = runs[len(runs) // 2]
        fam = {"gemm": 0.0, "attention": 0.0, "norm": 0.0, "other": 0.0}
        cnt = dict.fromkeys(fam, 0)
        t0 = run[0].time_range.start
        prev_end = t0
        for e in run:
            name = e.name
            k = "gemm" if re.search(r"gemm_tf32|conv3x3_halo|conv_smallk|gemv|splitk_reduce", name) else \
                "attention" if "attn" in name else "norm" if "norm" in name else "other"
            exposed = max(e.time_range.end - max(prev_end, e.time_range.start), 0.0) + max(e.time_range.start - prev_end, 0.0)
            prev_end = max(prev_end, e.time_range.end)
            fam[k] += exposed
            cnt[k] += 1
        span = prev_end - t0
        return {"kernels": len(run), "span_ms": span * 1e-3, "exposed_ms": {k: v * 1e-3 for k, v in fam.items()},
                "kernels_per_family": cnt,
                "gemm_tflops_exposed": gemm_flops / (fam["gemm"] * 1e-6) / 1e12 if fam["gemm"] > 0 else None,
                "note": "one replay of the captured step graph under CUPTI tracing; exposed = end - max(previous end, start); "
                        "the gemm family here includes the GEMV / small-K convolution / split-K reduce kernels"}
    except Exception as e:  # profiler unavailable: the block is optional
        return {"error": f"{type(e).__name__}: {e}"}


def run_unet20(args):
    from tsd_b200.api import Diffusion
    from tsd_b200.pipeline import Pipeline
    from tsd_b200.sampler import DDPMSampler, get_time_embedding
    torch, tdist, ctx, rank, world, local = setup_gpu(args)
    dev = torch.device("cuda", local)
    K, Wm = args.steps, args.warmup
    m = Diffusion(ctx, SIDE, SIDE, max_batch=1)
    m.init_random(1234)

    # ---- synthetic inputs (seeded per sample index = rank) -------------------------------------
    lat0, noise = tdist.sample_inputs(1234, rank, SIDE, LOOP_STEPS)
    context = np.random.default_rng(99).standard_normal((1, CTX_LEN, CTX_DIM), dtype=np.float32) if rank == 0 else None
    context = tdist.broadcast_context(context, (1, CTX_LEN, CTX_DIM))      # NCCL broadcast, once per prompt
    sm_ = DDPMSampler()
    sm_.set_inference_timesteps(LOOP_STEPS)
    ts = sm_.timesteps
    temb = np.stack([get_time_embedding(float(t)) for t in ts]).astype(np.float32)
    coef = sm_.coefficient_table()
    n_lat = 4 * SIDE * SIDE

    d_lat = torch.from_numpy(lat0.copy()).to(dev).reshape(1, 4, SIDE, SIDE).contiguous()
    d_lat0 = d_lat.clone()
    d_eps = torch.empty_like(d_lat)
    d_ctx = torch.from_numpy(context).to(dev).contiguous()
    d_temb = torch.from_numpy(temb).to(dev).contiguous()
    d_noise = torch.from_numpy(noise).to(dev).contiguous()
    torch.cuda.synchronize()

    def dev_step(i):
        j = i % LOOP_STEPS
        if j == 0:           # new image: reset the latent (64 KiB D2D, inside the timed region)
            ctx.synchronize()
            d_lat.copy_(d_lat0)
            torch.cuda.synchronize()
        m.forward_dev(d_lat.data_ptr(), None, 1, d_temb[j].data_ptr(), 1, 1, d_eps.data_ptr())
        ctx.sampler_step_dev(d_lat.data_ptr(), d_eps.data_ptr(), None, 1.0,
                             d_noise[j].data_ptr() if ts[j] > 0 else None, coef[j], n_lat, d_lat.data_ptr())

    # context + hoisted K/V projections: once per prompt, before the loop (not a timed step)
    m.forward_dev(d_lat.data_ptr(), d_ctx.data_ptr(), 1, d_temb[0].data_ptr(), 1, 1, d_eps.data_ptr())
    ctx.synchronize()

    def sync_all():
        ctx.synchronize()
        torch.cuda.synchronize()
        tdist.barrier()

    # ---- device-resident arm ---------------------------------------------------------------------
    for i in range(max(Wm, 3)):
        dev_step(i)
    sync_all()
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launch_count()
    ctx.timer_start()
    for i in range(K):
        dev_step(i)
    ms_dev = ctx.timer_stop()
    launches = ctx.launch_count() - l0
    sync_all()
    ms_dev = tdist.max_over_ranks(ms_dev)
    finite = bool(torch.isfinite(d_lat).all().item())

    # ---- end-to-end arm: host-buffer C ABI, pinned buffers, H2D/D2H every step ------------------
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    h_lat = pin(lat0.reshape(1, 4, SIDE, SIDE))
    h_lat0 = h_lat.clone().pin_memory()
    h_out = torch.empty_like(h_lat).pin_memory()
    h_ctx = pin(context)
    h_temb = pin(temb)
    h_noise = pin(noise)
    L = ctx.L

    def e2e_step(i):
        # one iteration of the reference loop through tsd_diffusion_step: H2D latents/time/noise, Diffusion.forward,
        # DDPMSampler.step, D2H latents; the context pointer is passed every call (its projections are reused while the
        # bytes are unchanged - the library hashes the 236 KB on the host every call, inside this timed region)
        j = i % LOOP_STEPS
        if j == 0:
            h_lat.copy_(h_lat0)
        c = [float(v) for v in coef[j]]
        ctx._ck(L.tsd_diffusion_step(m.m, h_lat.data_ptr(), h_ctx.data_ptr(), 1, h_temb[j].data_ptr(),
                                     h_noise[j].data_ptr() if ts[j] > 0 else None, 0, 1.0, *c, 1, h_out.data_ptr()))
        h_lat.copy_(h_out)

    for i in range(max(Wm, 3)):
        e2e_step(i)
    sync_all()
    ctx.timer_start()
    t0 = time.perf_counter()
    for i in range(K):
        e2e_step(i)
    ms_e2e_dev = ctx.timer_stop()
    ms_e2e = max(ms_e2e_dev, (time.perf_counter() - t0) * 1e3)
    sync_all()
    ms_e2e = tdist.max_over_ranks(ms_e2e)
    clk = clocks.stop() if clocks else None
    h2d = 4 * (n_lat + 320 + n_lat)   # latents, time, noise (the context only when its bytes change)
    d2h = 4 * n_lat                   # new latents

    # ---- whole image (config[1] end to end: 20 steps + VAE decode through pipeline.generate) -----
    image = None
    if not args.no_image:
        pipe = Pipeline(ctx, image_size=8 * SIDE, max_images=1, cfg=False, seed=1234)
        pipe.generate(context, inference_steps=LOOP_STEPS, latents=lat0[None], noise=noise[:, None])
        sync_all()
        reps = 3
        ctx.timer_start()
        t0 = time.perf_counter()
        for _ in range(reps):
            img, _ = pipe.generate(context, inference_steps=LOOP_STEPS, latents=lat0[None], noise=noise[:, None])
        ms_img = max(ctx.timer_stop(), (time.perf_counter() - t0) * 1e3) / reps
        sync_all()
        ms_img = tdist.max_over_ranks(ms_img)
        image = {"images_per_s": world * 1000.0 / ms_img, "ms_per_image": ms_img, "steps": LOOP_STEPS,
                 "includes": "H2D latents/noise/context, 20 UNet steps + sampler, VAE decode, rescale, D2H 3x512x512 image",
                 "finite": bool(np.isfinite(img).all())}
        pipe.diffusion.close()
        pipe.decoder.close()

    # ---- roofline of the dominant kernel (rank 0): per-launch CUDA events, eager replay of one step
    roof = None
    if rank == 0:
        m.forward_dev(d_lat.data_ptr(), d_ctx.data_ptr(), 1, d_temb[0].data_ptr(), 1, 1, d_eps.data_ptr())
        ctx.synchronize()
        acc = None
        reps = 3
        for _ in range(reps + 1):    # first pass warms the eager path
            fam = m.profile(d_lat.data_ptr(), None, 1, d_temb[3].data_ptr(), 1, 1, d_eps.data_ptr())
            if acc is None:
                acc = {k: dict(ms=0.0, flops=v["flops"], launches=v["launches"]) for k, v in fam.items()}
            else:
                for k, v in fam.items():
                    acc[k]["ms"] += v["ms"] / reps
        fam = acc
        burst, sustained, peak_src = tf32_peaks()
        g = fam["gemm"]
        ach = g["flops"] / (g["ms"] * 1e-3) / 1e12 if g["ms"] > 0 else 0.0
        a = fam["attention"]
        ach_a = a["flops"] / (a["ms"] * 1e-3) / 1e12 if a["ms"] > 0 else 0.0
        step_ach = (UNET_GFLOP - CTX_KV_GFLOP) * 1e9 * (K / (ms_dev * 1e-3)) / 1e12
        traffic, traffic_note = kernel_traffic()
        roof = {"bound": "tensor", "kernel": "gemm_tf32_kernel (tcgen05 kind::tf32 implicit-GEMM conv + linear)",
                "achieved": ach, "peak": burst, "unit": "TFLOP/s", "frac": ach / burst,
                "peak_sustained": sustained, "frac_of_sustained": ach / sustained,
                "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peak_src + "; frac uses the burst figure (the timed region is milliseconds long)",
                "own_cublas_tf32_measurement": own_tf32_measurement(),
                "launches_per_step": g["launches"], "flops_per_step": g["flops"], "ms_per_step_in_kernel": g["ms"],
                "note": "family times are per-launch CUDA events in an EAGER replay; the captured graph overlaps launches "
                        "(programmatic dependent launch), so their sum exceeds ms_per_step",
                "whole_step": {"achieved": step_ach, "frac": step_ach / burst, "frac_of_sustained": step_ach / sustained,
                               "flops_per_step": (UNET_GFLOP - CTX_KV_GFLOP) * 1e9},
                "attention": {"achieved": ach_a, "frac": ach_a / burst, "launches_per_step": a["launches"],
                              "ms_per_step_in_kernel": a["ms"], "flops_per_step": a["flops"]},
                "families_ms": {k: v["ms"] for k, v in fam.items()},
                "families_launches": {k: v["launches"] for k, v in fam.items()}}
        if not args.no_timeline:
            roof["graph_timeline"] = graph_timeline(lambda: dev_step(3), ctx.synchronize, g["flops"])
            tl = roof["graph_timeline"]
            if tl.get("gemm_tflops_exposed"):
                tl["gemm_frac_of_burst"] = tl["gemm_tflops_exposed"] / burst

    # ---- CPU baseline (rank 0, N = 1 only): ONE real UNet step on all host cores ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, n, secs, cores, cross = cpu_full_steps(30.0, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} full UNet step (oracle diffusion_forward(Ops('c32')) at the 4x64x64 latent + sampler step, "
                         f"oracle/ref_loops.c, {cores} OpenMP threads, {secs:.1f} s)", "sample_seconds": secs, **cross}

    if rank == 0:
        value = world * K / (ms_dev * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(Wm, 3),
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic", "config": base_config(world),
            "e2e": {"value": world * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K,
                    "api": "tsd_diffusion_step (host buffers, pinned): Diffusion.forward + DDPMSampler.step in one call"},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
            "image_e2e": image, "finite": finite,
        }
        print(json.dumps(line), flush=True)
    m.close()
    finish(ctx)


def run_cfg50(args):
    """BASELINE configs[2]: 50 steps, CFG 7.5, one image per rank; a bench step = one image end to end through
    Pipeline.generate (H2D latents/noise/contexts, 50 x UNet on [cond; uncond], sampler, VAE decode, D2H image)."""
    from tsd_b200.pipeline import Pipeline
    torch, tdist, ctx, rank, world, local = setup_gpu(args)
    K, Wm = args.steps, max(args.warmup, 3)
    steps = 50
    lat0, noise = tdist.sample_inputs(1234, rank, SIDE, steps)
    rng = np.random.default_rng(99)
    contexts = rng.standard_normal((2, CTX_LEN, CTX_DIM), dtype=np.float32) if rank == 0 else None
    contexts = tdist.broadcast_context(contexts, (2, CTX_LEN, CTX_DIM))    # cond + uncond, NCCL broadcast once
    pipe = Pipeline(ctx, image_size=8 * SIDE, max_images=1, cfg=True, seed=1234)

    def one_image(decode=True):
        return pipe.generate(contexts[0], contexts[1], cfg_scale=7.5, inference_steps=steps, latents=lat0[None],
                             noise=noise[:, None], decode=decode)

    def sync_all():
        ctx.synchronize()
        torch.cuda.synchronize()
        tdist.barrier()

    for _ in range(Wm):
        one_image()
    sync_all()
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launch_count()
    ctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(K):
        img, lat = one_image()
    ms_dev = ctx.timer_stop()
    ms_wall = (time.perf_counter() - t0) * 1e3
    launches = ctx.launch_count() - l0
    sync_all()
    ms_dev = tdist.max_over_ranks(ms_dev)
    ms_wall = tdist.max_over_ranks(max(ms_wall, ms_dev))
    # the loop alone (no decode): CFG steps/s
    ctx.timer_start()
    for _ in range(2):
        one_image(decode=False)
    ms_loop = tdist.max_over_ranks(ctx.timer_stop()) / 2
    clk = clocks.stop() if clocks else None
    n_lat = 4 * SIDE * SIDE
    if rank == 0:
        burst, sustained, peak_src = tf32_peaks()
        gflop_img = 100 * UNET_GFLOP + VAE_GFLOP
        ach = gflop_img * 1e9 * (K / (ms_dev * 1e-3)) / 1e12
        line = {"metric": "images_per_sec_512x512_50steps_cfg", "value": world * K / (ms_dev * 1e-3), "unit": "images/s",
                "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_dev / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
                "config": base_config(world, "cfg50"),
                "e2e": {"value": world * K / (ms_wall * 1e-3), "unit": "images/s", "ms_per_step": ms_wall / K,
                        "h2d_bytes_per_step": 4 * (n_lat + steps * n_lat + 2 * CTX_LEN * CTX_DIM + steps * 325),
                        "d2h_bytes_per_step": 4 * (3 * 512 * 512 + n_lat),
                        "api": "tsd_b200.pipeline.Pipeline.generate (tsd_generate_latents + tsd_decoder_forward, host buffers)"},
                "gpu_launches": int(launches), "clocks": clk,
                "roofline": {"bound": "tensor", "kernel": "whole image (100 UNet evaluations + VAE decode)", "achieved": ach,
                             "peak": burst, "unit": "TFLOP/s", "frac": ach / burst, "peak_sustained": sustained,
                             "frac_of_sustained": ach / sustained, "traffic": None, "peak_source": peak_src,
                             "flops_per_step": gflop_img * 1e9},
                "cfg_steps_per_s": world * steps * 1000.0 / ms_loop, "unet_evals_per_s": world * 2 * steps * 1000.0 / ms_loop,
                "ms_per_cfg_step": ms_loop / steps, "cpu_baseline": None,
                "finite": bool(np.isfinite(img).all() and np.isfinite(lat).all())}
        print(json.dumps(line), flush=True)
    pipe.close()
    finish(ctx)


def run_vae16(args):
    """BASELINE configs[3]: Decoder.forward (vae.mojo:221-250) at batch 16, the conv-roofline probe."""
    from tsd_b200.api import Decoder
    torch, tdist, ctx, rank, world, local = setup_gpu(args)
    dev = torch.device("cuda", local)
    K, Wm = args.steps, max(args.warmup, 3)
    B = args.batch
    dec = Decoder(ctx, SIDE, SIDE, max_batch=B)
    dec.init_random(1235)
    z = (np.random.default_rng(31 + rank).standard_normal((B, 4, SIDE, SIDE)) * 0.18215).astype(np.float32)
    d_z = torch.from_numpy(z).to(dev).contiguous()
    d_img = torch.empty((B, 3, 8 * SIDE, 8 * SIDE), device=dev, dtype=torch.float32)
    h_z = torch.from_numpy(z).pin_memory()
    h_img = torch.empty((B, 3, 8 * SIDE, 8 * SIDE), dtype=torch.float32).pin_memory()

    def sync_all():
        ctx.synchronize()
        torch.cuda.synchronize()
        tdist.barrier()

    for _ in range(Wm):
        dec.forward_dev(d_z.data_ptr(), B, True, d_img.data_ptr())
    sync_all()
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(K):
        dec.forward_dev(d_z.data_ptr(), B, True, d_img.data_ptr())
    ms_dev = ctx.timer_stop()
    launches = ctx.launch_count() - l0
    sync_all()
    ms_dev = tdist.max_over_ranks(ms_dev)
    finite = bool(torch.isfinite(d_img).all().item())
    L = ctx.L
    for _ in range(2):
        ctx._ck(L.tsd_decoder_forward(dec.m, h_z.data_ptr(), B, 1, h_img.data_ptr()))
    sync_all()
    t0 = time.perf_counter()
    for _ in range(K):
        ctx._ck(L.tsd_decoder_forward(dec.m, h_z.data_ptr(), B, 1, h_img.data_ptr()))
    ms_e2e = tdist.max_over_ranks((time.perf_counter() - t0) * 1e3)
    clk = clocks.stop() if clocks else None
    if rank == 0:
        burst, sustained, peak_src = tf32_peaks()
        ach = VAE_GFLOP * 1e9 * B * (K / (ms_dev * 1e-3)) / 1e12
        line = {"metric": "vae_decode_images_per_sec_512x512_bs16", "value": world * B * K / (ms_dev * 1e-3),
                "unit": "images/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_dev / K,
                "ms_per_image": ms_dev / K / B, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": DTYPE, "data": "synthetic", "config": dict(base_config(world, "vae16"), per_gpu_batch=B),
                "e2e": {"value": world * B * K / (ms_e2e * 1e-3), "unit": "images/s", "ms_per_step": ms_e2e / K,
                        "h2d_bytes_per_step": 4 * B * 4 * SIDE * SIDE, "d2h_bytes_per_step": 4 * B * 3 * 512 * 512,
                        "api": "tsd_decoder_forward (host buffers, pinned)"},
                "gpu_launches": int(launches), "clocks": clk,
                "roofline": {"bound": "tensor", "kernel": "Decoder.forward (98.3 % 3x3 implicit-GEMM conv, gemm_tf32_kernel)",
                             "achieved": ach, "peak": sustained, "unit": "TFLOP/s", "frac": ach / sustained,
                             "peak_burst": burst, "frac_of_burst": ach / burst, "traffic": None,
                             "peak_source": peak_src + "; frac uses the SUSTAINED figure (a batch-16 decode runs ~100 ms at the power cap)",
                             "flops_per_step": VAE_GFLOP * 1e9 * B},
                "cpu_baseline": None, "finite": finite}
        print(json.dumps(line), flush=True)
    dec.close()
    finish(ctx)


def run_attn(args):
    """BASELINE configs[4]: attention core (helpers/attention.mojo:46-62) h = 8, Tq = 4096, Tk in {4096, 77},
    d in {40, 80, 160}, both softmax axes (query axis = the reference's Softmax(dim=2), key axis = standard attention),
    device-timed, next to cuBLAS (torch.bmm + softmax + bmm, fp16 and TF32) and torch SDPA (fp16, key axis only)."""
    import ctypes as C
    torch, tdist, ctx, rank, world, local = setup_gpu(args)
    dev = torch.device("cuda", local)
    K, Wm = args.steps, max(args.warmup, 3)
    H, TQ = 8, 4096
    MUFU_PER_S = 16 * 148 * 1.965e9   # MUFU.EX2 lanes per clock per SM x SMs x clock

    def ours(tk, d, axis):
        ctx.set_option("softmax_axis", axis)
        ms = C.c_double()
        ctx._ck(ctx.L.tsd_bench_attention(ctx.h, H, TQ, tk, d, K, C.byref(ms)))
        return ms.value * 1e3

    def timed(fn):
        for _ in range(Wm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K * 1e3

    def torch_ref(tk, d, axis, dtype):
        torch.backends.cuda.matmul.allow_tf32 = True
        q = torch.randn(H, TQ, d, device=dev, dtype=dtype)
        k = torch.randn(H, tk, d, device=dev, dtype=dtype)
        v = torch.randn(H, tk, d, device=dev, dtype=dtype)
        sc = 1.0 / d ** 0.5
        out = {"cublas_us": timed(lambda: torch.bmm(torch.softmax(torch.bmm(q, k.transpose(1, 2)) * sc, dim=-1 if axis else -2), v))}
        if axis == 1 and dtype == torch.float16:
            out["sdpa_us"] = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q[None], k[None], v[None]))
        return out

    sweep = []
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launch_count()
    for tk in (4096, 77):
        for d in (40, 80, 160):
            for axis in (0, 1):
                us = ours(tk, d, axis)
                flops = 4.0 * H * TQ * tk * d
                row = {"tq": TQ, "tk": tk, "d": d, "softmax_axis": "query (reference)" if axis == 0 else "key (standard)",
                       "ours_us": us, "ours_tflops": flops / us / 1e6,
                       "exp_floor_us": 2.0 * H * TQ * tk / MUFU_PER_S * 1e6,
                       "io_mb_fp32": 4 * H * (2 * TQ + 2 * tk) * d / 1e6}
                f16 = torch_ref(tk, d, axis, torch.float16)
                f32 = torch_ref(tk, d, axis, torch.float32)
                row.update(cublas_fp16_us=f16["cublas_us"], cublas_tf32_us=f32["cublas_us"], sdpa_fp16_us=f16.get("sdpa_us"))
                row["speedup_vs_cublas_fp16"] = f16["cublas_us"] / us
                sweep.append(row)
    launches = ctx.launch_count() - l0
    ctx.set_option("softmax_axis", 0)
    clk = clocks.stop() if clocks else None
    if rank == 0:
        burst, sustained, peak_src = tf32_peaks()
        head = sweep[0]
        line = {"metric": "attention_core_tflops_T4096_d40", "value": head["ours_tflops"], "unit": "TFLOP/s", "n_gpus": 1,
                "steps": K, "warmup": Wm, "ms_per_step": head["ours_us"] / 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": DTYPE + "; baselines fp16 and TF32", "data": "synthetic",
                "config": base_config(1, "attn"),
                "e2e": {"value": head["ours_tflops"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                        "note": "kernel sweep: operands are generated on the device (tsd_bench_attention); the end-to-end "
                                "use of this kernel is the UNet step (--config unet20)"},
                "gpu_launches": int(launches), "clocks": clk,
                "roofline": {"bound": "tensor", "kernel": "attn2_kernel / attn_kernel (two passes: statistics, apply)",
                             "achieved": head["ours_tflops"], "peak": burst, "unit": "TFLOP/s", "frac": head["ours_tflops"] / burst,
                             "traffic": None, "peak_source": peak_src,
                             "note": "the op is MUFU.EX2-bound, not tensor-bound: 2 x h x Tq x Tk exponentials (two passes) at "
                                     "16 per clock per SM give exp_floor_us per case"},
                "sweep": sweep, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    finish(ctx)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="unet20", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=16, help="vae16: images per decode")
    ap.add_argument("--no-image", action="store_true", help="skip the whole-image (20 steps + VAE decode) leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-timeline", action="store_true", help="skip the CUPTI graph-replay timeline (roofline.graph_timeline)")
    args = ap.parse_args()
    defaults = {"unet20": (40, 5), "cfg50": (4, 3), "vae16": (5, 3), "attn": (20, 5)}
    if args.steps is None:
        args.steps = defaults[args.config][0]
    if args.warmup is None:
        args.warmup = defaults[args.config][1]
    if args.impl == "reference":
        run_reference(args)
    else:
        {"unet20": run_unet20, "cfg50": run_cfg50, "vae16": run_vae16, "attn": run_attn}[args.config](args)


if __name__ == "__main__":
    main()
