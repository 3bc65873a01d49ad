#!/usr/bin/env bash
# A/B of option sets: tests once, then short bench per setting.  Usage: tools/gpu_ab.sh "A=1 B=0" "A=0" ...
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
i=0
for setting in "$@"; do
  i=$((i+1))
  env $setting timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu > gpurun_out/bench_ab$i.json 2> gpurun_out/bench_ab$i.err
  echo "== [$setting] exit=$?"; tail -n 2 gpurun_out/bench_ab$i.err
  python - <<P
import json
d=json.load(open('gpurun_out/bench_ab$i.json'))
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), {k: round(v,3) for k,v in d['roofline']['families_ms'].items()}, 'img/s', round(d['image_e2e']['images_per_s'],2) if d.get('image_e2e') else None)
P
done
