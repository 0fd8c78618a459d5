cd /root/repo
for lib in libpomcp_b200.so libpomcp_b200_exp.so; do
echo "== $lib"
POMCP_B200_LIB=$PWD/pomdpy_b200/$lib timeout 300 python tools/episode_steplog.py 10 125000 8 1000000 2>&1 | grep "median\|move 3\|move 9"
POMCP_B200_LIB=$PWD/pomdpy_b200/$lib timeout 300 python tools/episode_steplog.py 14 2>&1 | grep "median\|move 4:\|move 11\|move 13"
done
POMCP_B200_LIB=$PWD/pomdpy_b200/libpomcp_b200.so python tools/episode_steplog.py 3 125000 8 1000000 > gpurun_out/plain.log 2>&1 && POMCP_B200_LIB=$PWD/pomdpy_b200/libpomcp_b200.so ncu --metrics gpu__time_duration.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pomcp_search_kernel -c 3 --csv --log-file gpurun_out/r02_local_mem_inline.csv python tools/episode_steplog.py 3 125000 8 1000000 > gpurun_out/ncu_local.log 2>&1
grep -c "" gpurun_out/r02_local_mem_inline.csv; grep "\"2\"" gpurun_out/r02_local_mem_inline.csv | awk -F'","' '{print $(NF-2), $NF}'
