#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ITERS=20 python tools/attn_lab.py > gpurun_out/r02b_attn_lab.txt 2>&1
TSD_ATTN_TRACE=1 ITERS=1 python tools/attn_lab.py 2>&1 | grep "attn2 trace" | head -24 > gpurun_out/r02b_attn_trace.txt
python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" 2>&1 | tail -15 > gpurun_out/r02b_pytest_attn.txt
python bench.py --steps 40 --warmup 5 --no-image --no-cpu > gpurun_out/r02b_bench_unet20.json 2> gpurun_out/r02b_bench_unet20.err
tail -3 gpurun_out/r02b_pytest_attn.txt
