cd /root/repo
python tools/lab/bisect_opts.py 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q -k "groupnorm or unet8 or unet64 or decoder or layernorm or encoder or clip_matches or loop" 2>&1 | tail -4
