cd /root/repo
timeout 300 python tools/episode_steplog.py 14 > gpurun_out/r2_episode.log 2>&1; echo "episode rc $?"
tail -19 gpurun_out/r2_episode.log | head -16
POMCP_B200_LIB=$PWD/pomdpy_b200/libpomcp_b200_mb3.so timeout 300 python tools/episode_steplog.py 14 > gpurun_out/r2_episode_mb3.log 2>&1; echo "episode mb3 rc $?"
tail -19 gpurun_out/r2_episode_mb3.log | head -16
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py tests/test_gpu_fuzz.py tests/test_gpu_debug_checks.py tests/test_gpu_tabular.py -q -x > gpurun_out/r2_gputests.log 2>&1; echo "gpu tests rc $?"
tail -5 gpurun_out/r2_gputests.log
timeout 600 python tools/gpu_quality.py map-15-15 10000 96 312,4,5000 > gpurun_out/r2_quality_15_small.txt 2>&1
timeout 600 python tools/gpu_quality.py map-15-15 100000 96 3125,4,50000 >> gpurun_out/r2_quality_15_small.txt 2>&1
timeout 600 python tools/gpu_quality.py map-15-15 1000 96 7,2,250 >> gpurun_out/r2_quality_15_small.txt 2>&1
cat gpurun_out/r2_quality_15_small.txt
