#!/usr/bin/env bash
# ncu --set full captures of the hot kernels at their UNet shapes (small reports: a few launches each)
mkdir -p gpurun_out
export TSD_OPT_autotune=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_kernel -c 2 -o gpurun_out/r01_ncu_attn python tools/prof_kernels.py > gpurun_out/prof_attn.log 2>&1; echo "attn exit=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32 --launch-skip 3 -c 1 -o gpurun_out/r01_ncu_gemm_conv320 python tools/prof_kernels.py > gpurun_out/prof_gemm.log 2>&1; echo "gemm exit=$?"
for f in r01_ncu_attn r01_ncu_gemm_conv320; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
  ncu -i gpurun_out/$f.ncu-rep --page details --csv > gpurun_out/$f.details.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv | tail
