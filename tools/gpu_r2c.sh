#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TSD_LIB=$PWD/stable-diffusion.mojo_b200/csrc/libtsd_b200_trace.so python tools/lab/gemm_trace.py > gpurun_out/r02c_gemm_trace.txt 2>&1
ITERS=20 python tools/prof_kernels.py > gpurun_out/r02c_kernel_timings.txt 2>&1
tail -30 gpurun_out/r02c_gemm_trace.txt
