cd /root/repo
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "unet or loop or generate" 2>&1 | tail -3
for o in "x=0" "gn_partial=0" "producer_stats=0" "norm_cluster=0"; do
echo "opt $o"
env TSD_OPT_$o timeout 600 python bench.py --no-image --no-cpu --steps 40 --warmup 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['roofline']['families_ms'], d['roofline']['families_launches'])"
done
