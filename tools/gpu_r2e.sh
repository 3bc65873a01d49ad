#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv or linear or matmul" 2>&1 | tail -6
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "unet8 or unet64 or blob or loop_with_cfg or decoder8" 2>&1 | tail -6
for v in 0 1; do
  TSD_OPT_splitk_cluster=$v python bench.py --steps 40 --warmup 5 --no-image --no-cpu > gpurun_out/r02e_bench_cluster$v.json 2> gpurun_out/r02e_bench_cluster$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r02e_bench_cluster$v.json')); print('cluster=$v', round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), {k:round(x,3) for k,x in d['roofline']['families_ms'].items()}, d['roofline']['families_launches'], d['gpu_launches'])"
done
