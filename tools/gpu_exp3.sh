#!/bin/bash
# GPU experiment 3: full parity suite (one-warp blocks, REDUX sums, on-device auto-ksize), tuning, full bench.
mkdir -p gpurun_out
bash tools/gpu_check.sh nobench > gpurun_out/check.log 2>&1
grep -E "^--- exit" gpurun_out/pytest.log | awk '{print $3}' | sort | uniq -c
grep -B2 -A25 -E "FAILED|Error|^E " gpurun_out/pytest.log | head -120
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
          " ".join(f"{k[:6]}={v}" for k, v in st.items()), "sum", round(sum(v for v in st.values() if v > 0), 3), flush=True)
except Exception as e:
    print(name, "FAILED", e, open(f"gpurun_out/bench_{name}.err").read()[-800:])
PY
}
run base X=1
run bps4 KR_EIG_BPS=4
BENCH_ARGS="--depth 8" run depth8 X=1
BENCH_ARGS="--depth 6" run depth6 X=1
BENCH_ARGS="--depth 1" run depth1 X=1
timeout 900 python bench.py --steps 20 --warmup 4 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "full bench exit $?"; tail -c 1500 gpurun_out/bench_full.json
