#!/usr/bin/env bash
# Round 2 evidence run: GPU tests, every bench config (+ reference arm), launch list of the step graph, ncu --set full
# captures of the kernels the north star names.  Everything lands in gpurun_out/ (kept < 64 MiB).
tag=${1:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.txt 2>&1
nproc > gpurun_out/nproc_$tag.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit=$?" | tee -a gpurun_out/pytest_$tag.log
grep -E "passed|failed" gpurun_out/pytest_$tag.log | tail -n 3
timeout 900 python bench.py --steps 40 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit=$?"
for cfgname in cfg50 vae16 attn; do
  timeout 900 python bench.py --config $cfgname > gpurun_out/bench_${cfgname}_$tag.json 2> gpurun_out/bench_${cfgname}_$tag.err; echo "bench $cfgname exit=$?"
done
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; echo "ref exit=$?"
# launch list of the step graph (per-launch device time; cold-cache and serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 250 -c 900 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-image --no-cpu --no-timeline > gpurun_out/ncu_bench_$tag.log 2>&1; echo "ncu list exit=$?"
cap() {  # name, kernel regex, launch-skip, count, group, extra ncu flags
  timeout 600 ncu --set full --clock-control none --import-source on $6 -k "regex:$2" --launch-skip $3 -c $4 \
    -o gpurun_out/${tag}_ncu_$1 python tools/prof_r02.py $5 > gpurun_out/prof_$1_$tag.log 2>&1; echo "ncu $1 exit=$?"
  ncu -i gpurun_out/${tag}_ncu_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_$1.raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/${tag}_ncu_$1.raw.csv gpurun_out/${tag}_ncu_$1 > /dev/null 2>&1
}
R="--profile-from-start off"
cap gemm_conv320 gemm_tf32 0 1 step "$R"
cap decoder_conv256 gemm_tf32 3 1 decoder
cap attn2 attn2_kernel 0 2 attn
cap cross_attn "attn_kernel" 0 2 attn
cap norm_apply_partial norm_apply_partial 0 1 step "$R"
cap norm_cluster norm_cluster 0 3 step "$R"
cap gemv_multi gemv_multi 0 1 step "$R"
# DRAM / L2->SM bytes and duration of EVERY launch of one warmed-up eager step
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --profile-from-start off --csv --log-file gpurun_out/${tag}_step_traffic.csv python tools/prof_r02.py step > /dev/null 2>&1; echo "ncu traffic exit=$?"
VERBOSE=1 timeout 300 python tools/lab/graph_timeline.py > gpurun_out/${tag}_graph_timeline.txt 2>&1
ITERS=20 timeout 300 python tools/prof_kernels.py > gpurun_out/kernel_timings_$tag.log 2>&1
ITERS=20 timeout 300 python tools/attn_lab.py > gpurun_out/attn_lab_$tag.txt 2>&1
python tools/step_launches.py gpurun_out/launches_$tag.csv gpurun_out/${tag}_launches_unet_step > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
du -sh gpurun_out
