#!/usr/bin/env bash
# tests + bench (+ optional launch list with NCU=1).  Usage: tools/gpu_tb.sh tag
tag=${1:-tb}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit=$?"
tail -n 3 gpurun_out/bench_$tag.err; python - <<P
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['families_ms'], 'frac', d['roofline']['frac'], 'whole', d['roofline']['whole_step']['frac'], 'img/s', d['image_e2e']['images_per_s'], 'launches', d['gpu_launches'])
P
if [ "${NCU:-0}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-image --no-cpu > gpurun_out/ncu_bench_$tag.log 2>&1; echo "ncu exit=$?"
fi
