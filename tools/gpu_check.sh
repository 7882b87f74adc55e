#!/bin/bash
# Run on the GPU box (under gpurun): every GPU parity test in its own process with a
# time limit (a hung kernel then costs one test, not the whole call), then a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
: > gpurun_out/pytest.log
TESTS=$(python - <<'PY'
import re
src = open("tests/test_gpu_parity.py").read()
print(" ".join(re.findall(r"^def (test_\w+)", src, re.M)))
PY
)
for t in $TESTS; do
  echo "=== $t" >> gpurun_out/pytest.log
  timeout 420 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "$t" -x --tb=short >> gpurun_out/pytest.log 2>&1
  echo "--- exit $? ($t)" >> gpurun_out/pytest.log
done
grep -E "^(===|---)|passed|failed|Error|error|assert" gpurun_out/pytest.log | tail -120
if [ "$1" != "nobench" ]; then
  timeout 600 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
  echo "bench exit $?"; tail -c 6000 gpurun_out/bench.log; tail -20 gpurun_out/bench.err
fi
