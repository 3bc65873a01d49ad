#!/usr/bin/env bash
# After `gpurun -- bash tools/gpu_round2.sh r02`: copy the condensed artefacts of the round into profiles/.
tag=${1:-r02}
set -e
cd "$(dirname "$0")/.."
g=gpurun_out
cp $g/bench_$tag.json profiles/${tag}_bench_unet20.json
for c in cfg50 vae16 attn; do cp $g/bench_${c}_$tag.json profiles/${tag}_bench_$c.json; done
cp $g/bench_ref_$tag.json profiles/${tag}_bench_ref.json
cp $g/pytest_$tag.log profiles/${tag}_pytest_gpu.txt
cp $g/kernel_timings_$tag.log profiles/${tag}_kernel_timings.log
cp $g/attn_lab_$tag.txt profiles/${tag}_attn_lab.txt
cp $g/${tag}_graph_timeline.txt profiles/${tag}_graph_timeline.txt
cp $g/${tag}_launches_unet_step.csv $g/${tag}_launches_unet_step_summary.txt profiles/
for k in gemm_conv320 decoder_conv256 attn2 cross_attn norm_apply_partial norm_cluster gemv_multi; do
  cp $g/${tag}_ncu_$k.json $g/${tag}_ncu_$k.txt profiles/
done
python tools/step_traffic.py $g/${tag}_step_traffic.csv profiles/${tag}_step_traffic.csv profiles/${tag}_kernel_traffic.json
ls profiles | grep "^${tag}_" | wc -l
