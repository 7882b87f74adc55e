#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_check.sh nobench > gpurun_out/check.log 2>&1
grep -E "^--- exit" gpurun_out/pytest.log | awk '{print $3}' | sort | uniq -c
grep -B2 -A25 -E "FAILED|Error|^E " gpurun_out/pytest.log | head -100
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
run nocache KR_LK_NOCACHE=1
BENCH_ARGS="--depth 1" run depth1 X=1
bash tools/ncu_profile.sh v11 > gpurun_out/ncu_profile.log 2>&1
tail -3 gpurun_out/ncu_profile.log
timeout 900 python bench.py --steps 20 --warmup 4 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "full bench exit $?"; head -c 1200 gpurun_out/bench_full.json
