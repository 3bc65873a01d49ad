cd /root/repo
run() { python bench.py --no-image --no-cpu --steps 40 --warmup 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['roofline']['families_ms'])"; }
for t in 256 128 64 512; do echo "target $t"; TSD_NORM_CLUSTER_CTAS=$t run; done
