#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02j_pytest_gpu.txt 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/r02j_pytest_gpu.txt
grep -E "passed|failed|rel_linf|error growth|CFG step" gpurun_out/r02j_pytest_gpu.txt | tail -40
python bench.py --config attn > gpurun_out/r02j_bench_attn.json 2> gpurun_out/r02j_bench_attn.err
python bench.py --config cfg50 > gpurun_out/r02j_bench_cfg50.json 2> gpurun_out/r02j_bench_cfg50.err
python bench.py --config vae16 > gpurun_out/r02j_bench_vae16.json 2> gpurun_out/r02j_bench_vae16.err
python bench.py > gpurun_out/r02j_bench_unet20.json 2> gpurun_out/r02j_bench_unet20.err
TSD_REF_BUDGET_S=60 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02j_bench_ref.json 2> gpurun_out/r02j_bench_ref.err
for f in attn cfg50 vae16 unet20 ref; do python -c "
import json; d=json.load(open('gpurun_out/r02j_bench_$f.json')); print('$f', d['metric'], round(d['value'],4), d['unit'], 'e2e', round(d['e2e']['value'],4), 'steps', d['steps'])"; done
