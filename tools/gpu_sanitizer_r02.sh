#!/usr/bin/env bash
# compute-sanitizer memcheck over the op tests that exercise this round's new kernels and the small models
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -m gpu -x -q \
  -k "groupnorm or layernorm or splitk_inside or conv2d_tile or linear or attention and not sweep" > gpurun_out/r02_sanitizer_ops.log 2>&1
echo "ops exit=$?" | tee -a gpurun_out/r02_sanitizer_ops.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_models.py -m gpu -x -q \
  -k "unet8_matches or unet8_execution or norm_affine or decoder8 or clip_matches" > gpurun_out/r02_sanitizer_models.log 2>&1
echo "models exit=$?" | tee -a gpurun_out/r02_sanitizer_models.log
tail -n 4 gpurun_out/r02_sanitizer_ops.log gpurun_out/r02_sanitizer_models.log
