#!/bin/bash
# Under gpurun (1 GPU): one ncu --set full capture of the kernels matching $1 (regex), tag $2.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$1" -c ${NCU_COUNT:-2} \
    -f -o gpurun_out/prof_$2 \
    python bench.py --steps 1 --warmup 0 --depth 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_one_$2.log 2>&1
ls -la gpurun_out/prof_$2.ncu-rep
