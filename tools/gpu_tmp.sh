cd /root/repo
python bench.py --no-image --no-cpu --steps 40 --warmup 5 2>gpurun_out/tmp_bench.err | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['roofline']['families_ms']); print(json.dumps(d['roofline'].get('graph_timeline'), indent=1))"
tail -3 gpurun_out/tmp_bench.err
