cd /root/repo
timeout 120 python tools/tc_gemm_check.py 4096 4096 4096 128 2>&1 | grep -v "first bad\|split ok" > gpurun_out/tc_check.log; echo "rc $?"
timeout 120 python tools/tc_gemm_check.py 4096 4096 4096 256 2>&1 | grep -v "first bad\|split ok" >> gpurun_out/tc_check.log; echo "rc $?"
timeout 120 python tools/tc_gemm_check.py 384 640 1056 128 2>&1 | grep -v "first bad\|split ok" >> gpurun_out/tc_check.log; echo "rc $?"
cat gpurun_out/tc_check.log
