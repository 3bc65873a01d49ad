cd /root/repo
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -s -k "norm_affine or decoder8 or clip_matches or encoder_matches" 2>&1 | tail -15
