cd /root/repo
timeout 40 tools/lab/l2bench
