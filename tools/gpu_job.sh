cd /root/repo
timeout 300 python tools/episode_steplog.py 14 > gpurun_out/r2_episode.log 2>&1; echo "episode rc $?"
tail -20 gpurun_out/r2_episode.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py tests/test_gpu_fuzz.py tests/test_gpu_debug_checks.py tests/test_gpu_tabular.py -q -x > gpurun_out/r2_gputests.log 2>&1; echo "gpu tests rc $?"
tail -5 gpurun_out/r2_gputests.log
timeout 900 python bench.py --steps 100 --warmup 3 --return_epochs 10 --no_cpu_baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc $?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['config']['fresh_tree_ms'], d['config']['search_ms'])"
