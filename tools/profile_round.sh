#!/usr/bin/env bash
# The commands behind profiles/r01_* (run on a B200 box from the repo root, e.g. through gpurun).
# ncu runs use `bench.py --no-graph`: ncu fails with LaunchFailed on CUDA-graph kernel nodes that take a
# CUtensorMap parameter (all TMA-fed kernels), so the identical kernels are launched eagerly for profiling.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
# bench lines (no profiler attached)
python bench.py --steps 20 --warmup 5 > gpurun_out/r01_bench_final.json || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r01_bench_reference_arm.json
for w in baseline sparse_attention wb2_64x32_ar_15f_4obs_4pred wb2_512x256_19f_ar; do
  python bench.py --workload "$w" --steps 10 --warmup 3 --cpu-baseline-seconds 10 --kernel-rows 8 | tail -1 > "gpurun_out/w_$w.json"
done
# the profiled command must have exited 0 without ncu first
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --profile-steps 0 > gpurun_out/r01_bench_nograph.json || exit 1
# launch list: ~2 steps of steady state (skip graph build + eager warm-ups)
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 -c 330 --csv \
    --log-file gpurun_out/r01_launches_attention_B64.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --profile-steps 0 > gpurun_out/ncu_l.log 2>&1
# full set on the dominant kernels; the .ncu-rep is exported to CSV on the box (gpurun_out/ is capped at 64 MiB)
ncu --set full --clock-control none -k regex:"gat_bwd|umma_dw_tma|umma_linear_tma|spmm_kernel|gat_alpha" \
    --launch-skip 430 -c 60 -o gpurun_out/r01_full -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --profile-steps 0 > gpurun_out/ncu_f.log 2>&1
ncu -i gpurun_out/r01_full.ncu-rep --page raw --csv > gpurun_out/r01_full_raw.csv
rm -f gpurun_out/r01_full.ncu-rep
# then, back home:
#   python tools/ncu_summary.py launches gpurun_out/r01_launches_attention_B64.csv > profiles/r01_launches_attention_B64_summary.csv
#   python tools/ncu_summary.py full gpurun_out/r01_full_raw.csv > profiles/r01_ncu_full_attention_B64.csv
#   python tools/ncu_roles.py <ncu --page source --csv export> [min_pct]     # stall samples per SASS instruction / warp role
