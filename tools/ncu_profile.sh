#!/bin/bash
# Under gpurun (1 GPU): launch list of a short bench run + one full capture of the heavy kernels.
mkdir -p gpurun_out
TAG=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --depth 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k "regex:k_eig_|k_exact_|k_laplacian|k_pyr_down|k_lk_roundtrip|k_minmax_mask|k_zncc|k_nms|k_mutual_info" -c ${NCU_COUNT:-12} \
    -f -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 0 --depth 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out/ | tail -8
grep -v "^==" gpurun_out/launches_${TAG}.csv | grep -E "k_|Kernel Name" | head -5
