#!/bin/bash
# GPU experiment: parity of the 4-pixel Laplacian + tuning sweep (quick bench runs).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x --tb=short \
  -k "laplacian or klt_match or smoke_entry or good_features or full_s2_scene" > gpurun_out/pytest_exp1.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_exp1.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 12 --warmup 4 ${BENCH_ARGS} > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{name}.json").read().strip().splitlines()[-1])
    st = {k: v["ms"] for k, v in d["stages"].items()}
    print(name, "ms/pair", round(d["ms_per_step"], 4), "pairs/s", round(d["scene_pairs_per_sec"], 1),
          "lap", st.get("laplacian_mon"), st.get("laplacian_ref"), "eig", st.get("corner_response"),
          "lk", st.get("lk_roundtrip"), "sum", round(sum(v for v in st.values() if v > 0), 3), flush=True)
except Exception as e:
    print(name, "FAILED", e, open(f"gpurun_out/bench_{name}.err").read()[-800:])
PY
}
run base X=1
run nolap4 KR_NO_LAP4=1
run bps5 KR_EIG_BPS=5
run bps5pad KR_EIG_BPS=5 KR_EIG_SMEM_PAD=8192
run seg64 KR_LAP4_SEG=64
run seg128 KR_LAP4_SEG=128
BENCH_ARGS="--depth 6" run depth6 X=1
BENCH_ARGS="--depth 8" run depth8 X=1
BENCH_ARGS="--depth 1" run depth1 X=1
