#!/usr/bin/env bash
# After `gpurun -- bash tools/gpu_round.sh rNN`: copy the condensed artefacts of that round into profiles/.
# Usage: tools/refresh_profiles.sh r01
tag=${1:-r01}
set -e
cp gpurun_out/bench_$tag.json profiles/${tag}_bench_final.json
cp gpurun_out/bench_ref_$tag.json profiles/${tag}_bench_ref_final.json
cp gpurun_out/kernel_timings_$tag.log profiles/${tag}_kernel_timings_final.log
cp gpurun_out/pytest_$tag.log profiles/${tag}_pytest_gpu_final.txt
python tools/step_launches.py gpurun_out/launches_$tag.csv profiles/${tag}_launches_unet_step_final | head -n 8
for k in gemm_conv320 attn norm; do python tools/ncu_summary.py gpurun_out/${tag}_ncu_$k.raw.csv profiles/${tag}_ncu_$k; done
