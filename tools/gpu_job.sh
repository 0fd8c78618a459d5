cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_gputests.log 2>&1; echo "gpu tests rc $?"
tail -12 gpurun_out/r2_gputests.log
