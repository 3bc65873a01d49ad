cd /root/repo
run() { python bench.py --no-image --no-cpu --steps 40 --warmup 5 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['roofline']['families_ms'], d['gpu_launches'])"; }
echo base; run
echo "cluster2, shipped plans"; TSD_OPT_splitk_cluster=2 run
export TSD_TUNE_DEFAULTS=0 TSD_TUNE_REPS=5 TSD_OPT_splitk_cluster=2
for i in 1 2; do
export TSD_TUNE_CACHE=gpurun_out/tune_r02_dsm$i.txt; rm -f $TSD_TUNE_CACHE
echo "cluster2 fresh tune $i"; run; run
done
