#!/usr/bin/env bash
# round 2, GPU call A: micro-benchmarks, attention v1/v2 A-B, all bench configs, then the GPU test suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02a_gpu.txt 2>&1
./tools/lab/ubench > gpurun_out/r02a_ubench.txt 2>&1
ITERS=20 python tools/attn_lab.py > gpurun_out/r02a_attn_lab.txt 2>&1
TSD_ATTN_TRACE=1 ITERS=1 python tools/attn_lab.py > gpurun_out/r02a_attn_trace.txt 2>&1
python bench.py --steps 40 --warmup 5 > gpurun_out/r02a_bench_unet20.json 2> gpurun_out/r02a_bench_unet20.err
TSD_OPT_attn_v2=0 python bench.py --steps 40 --warmup 5 --no-image --no-cpu > gpurun_out/r02a_bench_unet20_attn_v1.json 2> gpurun_out/r02a_bench_unet20_attn_v1.err
python bench.py --config attn > gpurun_out/r02a_bench_attn.json 2> gpurun_out/r02a_bench_attn.err
python bench.py --config cfg50 > gpurun_out/r02a_bench_cfg50.json 2> gpurun_out/r02a_bench_cfg50.err
python bench.py --config vae16 > gpurun_out/r02a_bench_vae16.json 2> gpurun_out/r02a_bench_vae16.err
gcc -O1 -o /tmp/abi_smoke tests/abi_smoke.c -ldl -lm && /tmp/abi_smoke stable-diffusion.mojo_b200/csrc/libtsd_b200.so expect-gpu > gpurun_out/r02a_abi_smoke.txt 2>&1
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -120 > gpurun_out/r02a_pytest_gpu.txt
tail -5 gpurun_out/r02a_pytest_gpu.txt
