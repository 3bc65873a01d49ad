#!/usr/bin/env bash
# parity probes (single CTAs, then pairs), in-kernel trace, tests
mkdir -p gpurun_out
for cg in 1 2; do
  for sec in gemm conv; do
    GEMM_CG=$cg timeout 300 python tools/gpu_probe.py $sec > gpurun_out/probe_${sec}_cg$cg.log 2>&1
    echo "== $sec cg=$cg exit=$?"; grep -E "BAD|FAILED|PROBE|Error|error|timed out" gpurun_out/probe_${sec}_cg$cg.log | head -20
  done
done
timeout 300 python tools/gemm_lab.py trace 2>&1 | grep -E "^\^\^|trace|FAILED|rror" | awk 'NR%6==1 || /\^\^/' 
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
