#!/usr/bin/env bash
# The commands behind profiles/r02_* (run on a B200 box from the repo root, e.g. through gpurun).
# ncu runs use `bench.py --no-graph`: ncu fails with LaunchFailed on CUDA-graph kernel nodes that take a
# CUtensorMap parameter (all TMA-fed kernels), so the identical kernels are launched eagerly for profiling.
# --min-timed-ms 0 keeps the timed region at exactly --steps steps (a run under ncu replays every kernel ~40x).
set -x
mkdir -p gpurun_out
W=${1:-wb2_512x256_19f_ar}
CMD="python bench.py --workload $W --no-workloads --no-cpu-baseline --no-graph --steps 2 --warmup 3 --min-timed-ms 0 --profile-steps 0"
STAGE=${2:-launches}      # one ncu invocation per box visit: run once with "launches", once with "full"
# the profiled command must have exited 0 without ncu first
$CMD > gpurun_out/r02_bench_nograph_$W.json 2> gpurun_out/plain.err || exit 1
if [ "$STAGE" = launches ]; then
  # launch list: ~2 steps of steady state (skip graph build, capture warm-ups and eager warm-ups)
  ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1400 -c 340 --csv \
      --log-file gpurun_out/r02_launches_$W.csv $CMD > gpurun_out/ncu_l.log 2>&1
else
  # full set on the dominant kernels; the .ncu-rep is exported to CSV on the box (gpurun_out/ is capped at 64 MiB)
  ncu --set full --import-source on --clock-control none \
      -k regex:"ws_kernel|umma_dw_ts|umma_dw_tma|umma_linear_ts|umma_linear_tma|spmm_heavy|gat_alpha_plan" \
      --launch-skip 400 -c 56 -o gpurun_out/r02_full_$W -f $CMD > gpurun_out/ncu_f.log 2>&1
  ncu -i gpurun_out/r02_full_$W.ncu-rep --page raw --csv > gpurun_out/r02_full_raw_$W.csv
  rm -f gpurun_out/r02_full_$W.ncu-rep
fi
# then, back home:
#   python tools/ncu_summary.py launches gpurun_out/r02_launches_$W.csv > profiles/r02_launches_${W}_summary.csv
#   python tools/ncu_summary.py full gpurun_out/r02_full_raw_$W.csv > profiles/r02_ncu_full_$W.csv
#   python tools/ncu_hot.py <x.ncu-rep> [min_pct]      # headline metrics + hottest SASS lines of one capture
