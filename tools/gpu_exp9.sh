#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --quick --steps 16 --warmup 4 ${BENCH_ARGS} > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{name}.json").read().strip().splitlines()[-1])
    print(name, "ms/pair", round(d["ms_per_step"], 4), "pairs/s", round(d["scene_pairs_per_sec"], 1), d["pipeline"], flush=True)
except Exception as e:
    print(name, "FAILED", e, open(f"gpurun_out/bench_{name}.err").read()[-800:])
PY
}
nproc; uptime
for i in 1 2 3 4 5 6 7 8; do
BENCH_ARGS="--depth 4" run d4_$i X=1
done
uptime
for i in 1 2 3 4; do
BENCH_ARGS="--depth 4" run d4nc_$i KR_LK_NOCACHE=1
done
