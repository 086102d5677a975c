#!/bin/bash
# Round-2 ncu evidence (run on a B200 box under gpurun; outputs in gpurun_out/).  Every profiled command first runs plain (&&).
#   tools/profile_r02.sh c3      launch list of the c3 step + --set full of its kernels
#   tools/profile_r02.sh other   --set full of the dominant kernels of c2 / c4 / c5 (solver + sampler update)
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
if [ "$1" = "c3" ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
  $CMD > gpurun_out/plain_c3.log 2>&1 && \
  $NCU --metrics gpu__time_duration.sum -s 60 -c 400 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launch_c3.log 2>&1
  echo "launch list rc=$?"
  $CMD > gpurun_out/plain_c3b.log 2>&1 && \
  $NCU --set full --import-source on -k regex:"npde_pair_grad|gram2_kernel|phi2_kernel|prep_x|prep_v|window_select" -s 30 -c 6 -o gpurun_out/prof_c3_r02 -f $CMD > gpurun_out/ncu_full_c3.log 2>&1
  echo "set full rc=$?"
else
  for wl in c4 c5 c2; do
    extra=""
    [ "$wl" = "c4" ] && extra="--tol loose"
    CMD="python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-graph $extra"
    $CMD > gpurun_out/plain_$wl.log 2>&1 && \
    $NCU --set full --import-source on -k regex:"dopri5_grad_kernel|sampler_kernel|hamcmc_kernel|npde_grad_kernel|npde_pair_grad" -s 8 -c 2 -o gpurun_out/prof_${wl}_r02 -f $CMD > gpurun_out/ncu_full_$wl.log 2>&1
    echo "$wl rc=$?"
  done
fi
