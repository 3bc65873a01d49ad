#!/usr/bin/env bash
mkdir -p gpurun_out
for cg in 1 2; do
  for sec in gemm conv; do
    TSD_OPT_autotune=0 GEMM_CG=$cg timeout 300 python tools/gpu_probe.py $sec > gpurun_out/probe_${sec}_cg$cg.log 2>&1
    echo "== $sec cg=$cg exit=$?"; grep -E "BAD|PROBE|timed out" gpurun_out/probe_${sec}_cg$cg.log | head -8
  done
done
for cg in 1 2; do for bn in 80 160; do for dbg in 0 1 6; do timeout 60 python tools/lab/roles_one.py $bn $cg $dbg 1 2>&1 | grep -E "us$|timed out" | head -2; done; done; done
bash tools/gpu_ab.sh "TSD_OPT_autotune=1" "TSD_OPT_autotune=0"
