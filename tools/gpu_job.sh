cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2_gputests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc $?"
python -c "
import json; s=open('gpurun_out/r02_bench_n1.json').read(); d=json.loads(s[s.index('{'):]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['search_ms'], d['config']['fresh_tree_ms'], d['mean_discounted_return'], d['return_stderr'])"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc $?"; cut -c1-200 gpurun_out/r02_bench_ref.json
python tools/gpu_quality.py map-15-15 1000000 192 125000,32,1000000 > gpurun_out/r2_quality_final.txt 2>&1
python tools/gpu_quality.py map-11-11 100000 192 12500,32,100000 >> gpurun_out/r2_quality_final.txt 2>&1; cat gpurun_out/r2_quality_final.txt
B="python bench.py --steps 20 --warmup 3 --return_epochs 1 --no_cpu_baseline --no_secondary"
timeout 600 $B > gpurun_out/plain.log 2>&1; echo "plain rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k pomcp_search_kernel -s 40 -c 14 -f -o gpurun_out/r02_prof_bench $B > gpurun_out/ncu_bench.log 2>&1; echo "ncu full rc $?"
ls -la gpurun_out/r02_prof_bench.ncu-rep
