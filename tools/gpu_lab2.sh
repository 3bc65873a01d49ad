#!/usr/bin/env bash
# Lab round: parity tests, feed experiments, per-shape sweep, bench.
tag=${1:-lab2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit=$?"
tail -n 5 gpurun_out/pytest_$tag.log
timeout 600 python tools/gemm_lab.py feed > gpurun_out/lab_feed_$tag.log 2>&1; echo "feed exit=$?"
grep -E "stages=max" gpurun_out/lab_feed_$tag.log
timeout 600 python tools/gemm_lab.py shapes > gpurun_out/lab_shapes_$tag.log 2>&1; echo "shapes exit=$?"
grep -E "best|heuristic|FAILED" gpurun_out/lab_shapes_$tag.log
timeout 600 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit=$?"
tail -n 3 gpurun_out/bench_$tag.err; cut -c1-2500 gpurun_out/bench_$tag.json
