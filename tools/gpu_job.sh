cd /root/repo
PREF=1 timeout 300 python tools/gpu_quality.py map-15-15 1000000 96 125000,32,1000000 > gpurun_out/r2_quality_pref.txt 2>&1
PREF=2 timeout 300 python tools/gpu_quality.py map-15-15 1000000 96 125000,32,1000000 >> gpurun_out/r2_quality_pref.txt 2>&1
PREF=2 timeout 300 python tools/gpu_quality.py map-15-15 100000 96 12500,32,100000 3125,4,50000 >> gpurun_out/r2_quality_pref.txt 2>&1
PREF=2 timeout 300 python tools/gpu_quality.py map-11-11 100000 96 12500,32,100000 >> gpurun_out/r2_quality_pref.txt 2>&1
tail -c 3000 gpurun_out/r2_quality_pref.txt
