#!/usr/bin/env bash
# Two-GPU leg of the evidence run (gpurun --gpus 2): torchrun benches (unet20, cfg50, reference arm) and the C-ABI
# multi-GPU check (tests/dist_c_abi_check.py through pytest).
tag=${1:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/bench_2gpu_$tag.json 2> gpurun_out/bench_2gpu_$tag.err; echo "bench 2gpu exit=$?"
timeout 900 $TR bench.py --gpus 2 --config cfg50 > gpurun_out/bench_cfg50_2gpu_$tag.json 2> gpurun_out/bench_cfg50_2gpu_$tag.err; echo "cfg50 2gpu exit=$?"
timeout 900 $TR bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_2gpu_$tag.json 2> gpurun_out/bench_ref_2gpu_$tag.err; echo "ref 2gpu exit=$?"
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q -s -k "dist_c_abi" > gpurun_out/pytest_2gpu_$tag.log 2>&1; echo "pytest dist exit=$?"
tail -n 3 gpurun_out/pytest_2gpu_$tag.log
grep -h '^{' gpurun_out/bench_2gpu_$tag.json gpurun_out/bench_cfg50_2gpu_$tag.json gpurun_out/bench_ref_2gpu_$tag.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d.get('impl', 'ours'), d['metric'], d['value'], d['unit'], 'n_gpus', d['n_gpus'])"
