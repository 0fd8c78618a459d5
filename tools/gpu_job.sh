cd /root/repo
timeout 300 python tools/debug_pref.py > gpurun_out/r2_debug_pref.log 2>&1; echo "debug rc $?"
tail -40 gpurun_out/r2_debug_pref.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -q > gpurun_out/r2_parity.log 2>&1; echo "parity rc $?"
tail -15 gpurun_out/r2_parity.log
timeout 300 python tools/episode_steplog.py 14 > gpurun_out/r2_episode.log 2>&1; echo "episode rc $?"
tail -24 gpurun_out/r2_episode.log
POMCP_B200_LIB=$PWD/pomdpy_b200/libpomcp_b200_tsec.so timeout 300 python tools/episode_steplog.py 1 > gpurun_out/r2_tsec_fresh.log 2>&1
tail -8 gpurun_out/r2_tsec_fresh.log
POMCP_B200_LIB=$PWD/pomdpy_b200/libpomcp_b200_tsec.so timeout 300 python tools/episode_steplog.py 10 > gpurun_out/r2_tsec_ep.log 2>&1
tail -8 gpurun_out/r2_tsec_ep.log
