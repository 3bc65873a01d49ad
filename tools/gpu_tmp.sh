cd /root/repo
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv or linear or matmul or gemm" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "unet or decoder8 or encoder_matches" 2>&1 | tail -3
run() { python bench.py --no-image --no-cpu --steps 40 --warmup 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['roofline']['families_ms'])"; }
runv() { python bench.py --config vae16 --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('vae16', d['value'], d['clocks'])"; }
run; TSD_OPT_gemm_kmerge=0 run; run; runv; TSD_OPT_gemm_kmerge=0 runv
