cd /root/repo
python tools/lab/norm_lab.py 2>&1 | head -3
TSD_NORM_TRACE=1 python tools/lab/norm_lab.py 2>&1 | grep "fold detail\|trace block" | head -6
