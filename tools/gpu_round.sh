#!/usr/bin/env bash
# One GPU round: parity tests, bench line, ncu launch list.  Usage: tools/gpu_round.sh [tag]
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.txt 2>&1
nproc > gpurun_out/nproc_$tag.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/pytest_$tag.log
tail -n 25 gpurun_out/pytest_$tag.log
timeout 900 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit=$?"
tail -n 5 gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>&1; echo "ref exit=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 1 --no-image --no-cpu > gpurun_out/ncu_bench_$tag.log 2>&1; echo "ncu exit=$?"
tail -n 3 gpurun_out/ncu_bench_$tag.log
