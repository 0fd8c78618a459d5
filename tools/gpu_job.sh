cd /root/repo
timeout 300 python tools/episode_steplog.py 14 > gpurun_out/r2_episode.log 2>&1; echo "episode rc $?"
tail -19 gpurun_out/r2_episode.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -q -x > gpurun_out/r2_gputests.log 2>&1; echo "gpu tests rc $?"
tail -4 gpurun_out/r2_gputests.log
