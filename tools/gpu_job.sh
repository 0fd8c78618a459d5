cd /root/repo
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 3 --return_epochs_multi 16 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc $?"
tail -2 gpurun_out/r2_bench_n$N.err
python -c "
import json,sys; d=json.load(open('gpurun_out/r2_bench_n$N.json')); print({k:d[k] for k in ('value','ms_per_step','scaling','mean_discounted_return','return_stderr')}, d['e2e'], d['other_scaling'], d['config']['merge_check'], d['config']['search_ms'], d['config']['fresh_tree_ms'])"
