#!/usr/bin/env bash
# Lab round 1: feed experiments + per-shape sweep + ncu --set full of representative kernels.
mkdir -p gpurun_out
timeout 600 python tools/gemm_lab.py feed > gpurun_out/lab_feed.log 2>&1; echo "feed exit=$?"
tail -n 80 gpurun_out/lab_feed.log
timeout 600 python tools/gemm_lab.py shapes > gpurun_out/lab_shapes.log 2>&1; echo "shapes exit=$?"
grep -E "best|heuristic|FAILED" gpurun_out/lab_shapes.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 --launch-skip 3 --launch-count 1 \
  -o gpurun_out/ncu_gemm_conv320 python tools/gemm_lab.py ncu > gpurun_out/ncu_gemm1.log 2>&1; echo "ncu1 exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 --launch-skip 7 --launch-count 9 \
  -o gpurun_out/ncu_gemm_more python tools/gemm_lab.py ncu > gpurun_out/ncu_gemm2.log 2>&1; echo "ncu2 exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_kernel|norm_fused" --launch-skip 91 --launch-count 40 \
  -o gpurun_out/ncu_attn_norm python bench.py --steps 1 --warmup 1 --no-image --no-cpu > gpurun_out/ncu_an.log 2>&1; echo "ncu3 exit=$?"
for f in ncu_gemm_conv320 ncu_gemm_more ncu_attn_norm; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -n 20
