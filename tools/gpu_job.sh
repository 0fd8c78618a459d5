cd /root/repo
timeout 1800 python tools/gpu_quality.py map-15-15 1000000 192 31250,4,1000000 62500,8,1000000 62500,4,1000000 125000,8,1000000 31250,4,500000 > gpurun_out/r2_quality_onegen192.txt 2>&1
cat gpurun_out/r2_quality_onegen192.txt
