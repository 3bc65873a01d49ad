cd /root/repo
TSD_LIB=$PWD/stable-diffusion.mojo_b200/csrc/libtsd_b200_trace.so timeout 300 python tools/lab/gemm_trace.py 2>&1 | tail -40
