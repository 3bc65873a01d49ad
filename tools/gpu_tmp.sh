cd /root/repo
nvidia-smi -L
python -m pytest tests/test_gpu_models.py -m gpu -x -q -s -k "dist_c_abi" 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 40 --warmup 5 --no-image --no-cpu > gpurun_out/r02k_bench_2gpu.json 2> gpurun_out/r02k_bench_2gpu.err; tail -2 gpurun_out/r02k_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --config cfg50 > gpurun_out/r02k_bench_cfg50_2gpu.json 2> gpurun_out/r02k_bench_cfg50_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r02k_bench_ref_2gpu.json 2> gpurun_out/r02k_bench_ref_2gpu.err
for f in 2gpu cfg50_2gpu ref_2gpu; do python -c "
import json; d=json.load(open('gpurun_out/r02k_bench_$f.json')); print('$f', d['metric'], round(d['value'],3), d['unit'], 'n_gpus', d['n_gpus'], 'e2e', round(d['e2e']['value'],3), (d.get('cpu_baseline') or {}).get('cores'))"; done
