#!/usr/bin/env bash
# quick GPU iteration: parity tests (all, not -x) + a short bench.  Usage: tools/gpu_quick.sh tag [pytest -k expr]
tag=${1:-q}; kexpr=${2:-}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then
  timeout 1200 python -m pytest tests -m gpu -q -s -k "$kexpr" > gpurun_out/pytest_$tag.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$tag.log 2>&1
fi
echo "pytest exit=$?" | tee -a gpurun_out/pytest_$tag.log
grep -E "passed|failed|rel_linf|FAILED|Error|error" gpurun_out/pytest_$tag.log | tail -n 40
timeout 600 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit=$?"
tail -n 5 gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json
if [ "${NCU:-0}" = "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-image --no-cpu > gpurun_out/ncu_bench_$tag.log 2>&1; echo "ncu exit=$?"
fi
