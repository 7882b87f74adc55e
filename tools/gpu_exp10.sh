#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" KR_TRACE_UNITS=1 timeout 300 python bench.py --quick --steps 16 --warmup 4 ${BENCH_ARGS} > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{name}.json").read().strip().splitlines()[-1])
    print(name, "ms/pair", round(d["ms_per_step"], 4), "pairs/s", round(d["scene_pairs_per_sec"], 1), d["pipeline"], flush=True)
    if d["ms_per_step"] > 1.8:
        print(open(f"gpurun_out/bench_{name}.err").read()[-1800:])
except Exception as e:
    print(name, "FAILED", e, open(f"gpurun_out/bench_{name}.err").read()[-800:])
PY
}
sleep 45
for i in 1 2 3 4 5 6; do
BENCH_ARGS="--depth 4" run d4_$i X=1
done
