#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python tools/gemm_lab.py shapes > gpurun_out/lab_shapes4.log 2>&1; echo "shapes exit=$?"
grep -E "best|heuristic|FAILED" gpurun_out/lab_shapes4.log
timeout 600 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_lab4.json 2> gpurun_out/bench_lab4.err; echo "bench exit=$?"
tail -n 3 gpurun_out/bench_lab4.err; python - <<'P'
import json
d=json.load(open('gpurun_out/bench_lab4.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['families_ms'], d['roofline']['frac'], d['image_e2e'])
P
