#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/lab/norm_lab.py > gpurun_out/r02g_norm_lab.txt 2>&1
TSD_NORM_TRACE=1 python tools/lab/norm_lab.py 2>&1 | grep "fold detail\|trace block" | head -8 >> gpurun_out/r02g_norm_lab.txt
cat gpurun_out/r02g_norm_lab.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "groupnorm or unet8 or unet64 or decoder8 or layernorm or encoder or clip_matches" 2>&1 | tail -4
python bench.py --steps 40 --warmup 5 --no-image --no-cpu > gpurun_out/r02g_bench_unet20.json 2> gpurun_out/r02g_bench_unet20.err
python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_unet20.json')); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), {k:round(x,3) for k,x in d['roofline']['families_ms'].items()}, d['gpu_launches'])"
