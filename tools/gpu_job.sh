cd /root/repo
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --return_epochs_multi 16 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench n$N rc $?"
python -c "
import json,sys; s=open('gpurun_out/r02_bench_n$N.json').read(); d=json.loads(s[s.index('{'):]); print({k:d[k] for k in ('value','ms_per_step','scaling','mean_discounted_return','return_stderr')}, d['e2e']['value'], d['other_scaling']['value'], d['other_scaling']['ms_per_step'], d['config']['search_ms'], d['config']['fresh_tree_ms'], d['config']['search_kernel_ms_before_merge_median'], d['config']['merge_wait_for_slowest_peer_ms_median'], d['config']['merge_check'])"
