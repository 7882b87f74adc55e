#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -x --tb=short \
  -k "sweep or pyr or klt_match or klt_tracker or smoke_entry or full_s2 or match_many or large_shift" > gpurun_out/pytest_exp14.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_exp14.log; grep -E "^E " gpurun_out/pytest_exp14.log | head -20
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 20 --warmup 4 ${BENCH_ARGS} > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{name}.json").read().strip().splitlines()[-1])
    st = {k: v["ms"] for k, v in d["stages"].items()}
    print(name, "ms/pair", round(d["ms_per_step"], 4), "pairs/s", round(d["scene_pairs_per_sec"], 1), d["pipeline"]["max_unit_gap_ms"],
          " ".join(f"{k[:6]}={v}" for k, v in st.items()), "sum", round(sum(v for v in st.values() if v > 0), 3), flush=True)
except Exception as e:
    print(name, "FAILED", e, open(f"gpurun_out/bench_{name}.err").read()[-800:])
PY
}
run base X=1
run base2 X=1
