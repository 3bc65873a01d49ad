cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q -s -k "norm_affine or unet64_full or decoder64 or checkpoint_round or unet8_matches" 2>&1 | grep -E "passed|failed|rel_linf|Error|assert" | tail -14
python bench.py --config vae16 > gpurun_out/r02k_bench_vae16.json 2> gpurun_out/r02k_bench_vae16.err
python -c "
import json; d=json.load(open('gpurun_out/r02k_bench_vae16.json')); print('vae16', round(d['value'],1), 'img/s', round(d['ms_per_image'],3), 'ms/img', 'frac', round(d['roofline']['frac'],3))"
python tools/time_decoder.py 2>&1 | tail -4
