#!/usr/bin/env bash
# One full GPU round: parity tests, bench (+ reference arm), launch list of the step graph, ncu --set full
# captures of the dominant kernels.  Everything lands in gpurun_out/ (kept small: < 64 MiB).
# Usage: tools/gpu_round.sh [tag]
tag=${1:-r01}
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
export TSD_TUNE_CACHE=$PWD/gpurun_out/tune_cache_$tag.txt
rm -f "$TSD_TUNE_CACHE"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.txt 2>&1
nproc > gpurun_out/nproc_$tag.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/pytest_$tag.log
grep -E "passed|failed|rel_linf" gpurun_out/pytest_$tag.log | tail -n 14
rm -f "$TSD_TUNE_CACHE"   # the tests tune many throw-away shapes: start the bench from a clean cache
timeout 900 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit=$?"
tail -n 3 gpurun_out/bench_$tag.err; cut -c1-400 gpurun_out/bench_$tag.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>&1; echo "ref exit=$?"
# launch list of the step graph: the tune cache is warm, so no tuning kernels are launched
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 250 -c 900 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-image --no-cpu > gpurun_out/ncu_bench_$tag.log 2>&1; echo "ncu list exit=$?"
# full captures of the dominant kernels at their UNet shapes (cost-model plan: deterministic launch order)
export TSD_OPT_autotune=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 --launch-skip 3 -c 1 \
  -o gpurun_out/${tag}_ncu_gemm_conv320 python tools/prof_kernels.py > gpurun_out/prof_gemm_$tag.log 2>&1; echo "ncu gemm exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -c 2 \
  -o gpurun_out/${tag}_ncu_attn python tools/prof_kernels.py > gpurun_out/prof_attn_$tag.log 2>&1; echo "ncu attn exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"norm_fused|norm_apply" -c 2 \
  -o gpurun_out/${tag}_ncu_norm python tools/prof_norm.py > gpurun_out/prof_norm_$tag.log 2>&1; echo "ncu norm exit=$?"
for f in ${tag}_ncu_gemm_conv320 ${tag}_ncu_attn ${tag}_ncu_norm; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ITERS=20 timeout 300 python tools/prof_kernels.py > gpurun_out/kernel_timings_$tag.log 2>&1
du -sh gpurun_out
# condensed artefacts for profiles/ (copy them over after reading the numbers):
python tools/step_launches.py gpurun_out/launches_$tag.csv gpurun_out/${tag}_launches_unet_step > /dev/null 2>&1
for f in ${tag}_ncu_gemm_conv320 ${tag}_ncu_attn ${tag}_ncu_norm; do
  python tools/ncu_summary.py gpurun_out/$f.raw.csv gpurun_out/$f > /dev/null 2>&1
done
rm -f gpurun_out/*.ncu-rep
