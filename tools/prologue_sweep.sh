for v in base coop0 coop2 coop12 coop32; do
  lib=/root/repo/pomdpy_b200/libpomcp_b200_$v.so; [ $v = base ] && lib=/root/repo/pomdpy_b200/libpomcp_b200.so
  POMCP_B200_LIB=$lib timeout 100 python tools/episode_steplog.py 80 125000 32 1000000 > gpurun_out/sweep_$v.log 2>&1
  awk -v c="$v" '/^move/ {t=$3+0; n++; s+=t; if (n>=37 && n<=79) {b++; sb+=t} if (n>=2 && n<=35) {h++; sh+=t}} /^phase totals/ {x=x" | "$0} END {printf "cfg %s: mean %.4f moves36-78 %.4f moves1-34 %.4f %s\n", c, s/n, sb/b, sh/h, x}' gpurun_out/sweep_$v.log
done
