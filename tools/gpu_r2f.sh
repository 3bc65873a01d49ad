#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/lab/splitk_lab.py > gpurun_out/r02f_splitk_lab.txt 2>&1
TRACE=1 TSD_LIB=$PWD/stable-diffusion.mojo_b200/csrc/libtsd_b200_trace.so python tools/lab/splitk_lab.py > gpurun_out/r02f_splitk_trace.txt 2>&1
cat gpurun_out/r02f_splitk_lab.txt; grep -v "chunk 0" gpurun_out/r02f_splitk_trace.txt | head -60
