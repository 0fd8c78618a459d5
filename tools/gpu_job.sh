cd /root/repo
timeout 600 python bench_vi.py > gpurun_out/r2_bench_vi.jsonl 2> gpurun_out/r2_bench_vi.err; echo "bench_vi rc $?"; sed -n 2p gpurun_out/r2_bench_vi.jsonl | python -c "
import json,sys; d=json.loads(sys.stdin.read()); t=d['tcgen05_tf32x3']; print(d['ms'], t['ms'], t['persistent_2_accumulator_sets'])"
timeout 90 python tools/tc_gemm_check.py 4096 4096 4096 persistent 2>&1 | grep "per GEMM\|variant" | cut -c1-120
timeout 90 python tools/tc_gemm_check.py 2048 2560 160 persistent 2>&1 | grep "per GEMM\|variant" | cut -c1-120
