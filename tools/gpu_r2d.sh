#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for nb in 296 592; do
  echo "== TSD_NORM_BLOCKS=$nb"
  TSD_NORM_BLOCKS=$nb python tools/lab/norm_lab.py 2>&1
done > gpurun_out/r02d_norm_lab.txt
TSD_NORM_TRACE=1 python tools/lab/norm_lab.py 2>&1 | grep "norm_apply_partial trace" | head -8 >> gpurun_out/r02d_norm_lab.txt
cat gpurun_out/r02d_norm_lab.txt
python -m pytest tests -m gpu -x -q -k "groupnorm or unet8 or unet64 or decoder8 or layernorm" 2>&1 | tail -4
python bench.py --steps 40 --warmup 5 --no-image --no-cpu > gpurun_out/r02d_bench_unet20.json 2> gpurun_out/r02d_bench_unet20.err
python -c "
import json; d=json.load(open('gpurun_out/r02d_bench_unet20.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['families_ms'])"
