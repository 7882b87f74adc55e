#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x --tb=short \
  -k "good_features or corner_modes or klt_match or klt_tracker or smoke_entry or full_s2 or match_many or sort_and_unlimited" > gpurun_out/pytest_exp8.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_exp8.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 16 --warmup 4 ${BENCH_ARGS} > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{name}.json").read().strip().splitlines()[-1])
    st = {k: v["ms"] for k, v in d["stages"].items()}
    print(name, "ms/pair", round(d["ms_per_step"], 4), "pairs/s", round(d["scene_pairs_per_sec"], 1), "lk", st["lk_roundtrip"], "nms", st["nms"], "rounds", d["stats_last"]["nms_rounds"], flush=True)
except Exception as e:
    print(name, "FAILED", e, open(f"gpurun_out/bench_{name}.err").read()[-800:])
PY
}
for i in 1 2 3 4 5 6; do
BENCH_ARGS="--depth 4" run d4_$i X=1
done
for i in 1 2 3; do
BENCH_ARGS="--depth 6" run d6_$i X=1
BENCH_ARGS="--depth 8" run d8_$i X=1
done
