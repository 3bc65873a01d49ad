#!/usr/bin/env bash
# Runs every probe section in its own process under a timeout; logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for sec in "$@"; do
  echo "=== $sec ===" 
  timeout 300 python tools/gpu_probe.py "$sec" > "gpurun_out/probe_$sec.log" 2>&1
  echo "exit=$?" >> "gpurun_out/probe_$sec.log"
  tail -n 60 "gpurun_out/probe_$sec.log"
done
