cd /root/repo
timeout 600 python -m pytest tests/test_vi.py -m gpu -x -q > gpurun_out/r2_vi_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2_vi_tests.log
timeout 600 python bench_vi.py > gpurun_out/r2_bench_vi.jsonl 2> gpurun_out/r2_bench_vi.err; echo "bench_vi rc $?"; sed -n 2p gpurun_out/r2_bench_vi.jsonl | cut -c1-2200
timeout 120 python tools/tc_gemm_check.py 4096 4096 4096 tile128 > gpurun_out/tc_check.log 2>&1; echo "rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vi_tf32x3 -s 1 -c 1 -f -o gpurun_out/r02_prof_tcgen05 python tools/tc_gemm_check.py 4096 4096 4096 tile128 > gpurun_out/ncu_tc.log 2>&1; echo "ncu rc $?"
