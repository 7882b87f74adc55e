#!/bin/bash
# Under gpurun (1 GPU): quick bench runs (resident throughput + stage times) for a list of
# environment settings, e.g.
#   tools/gpu_tune.sh base: lseg96:KR_LAP4_SEG=96 bps4:KR_EIG_BPS=4 pw4:KR_PYR_WARPS=4 nocache:KR_LK_NOCACHE=1
# Tunables: KR_LAP4_SEG (rows per Laplacian block), KR_NO_LAP4, KR_EIG_BPS (4|5), KR_EIG_SEG,
# KR_EIG_SMEM_PAD, KR_PYR_WARPS (1|2|4|8), KR_LK_NOCACHE, KR_NMS_PLAIN, KR_TRACE_UNITS.
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  env ${envs//,/ } X=1 timeout 300 python bench.py --quick --steps ${STEPS:-20} --warmup 4 ${BENCH_ARGS} \
      > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{name}.json").read().strip().splitlines()[-1])
    st = {k: v["ms"] for k, v in d["stages"].items()}
    print(name, "ms/pair", round(d["ms_per_step"], 4), "pairs/s", round(d["scene_pairs_per_sec"], 1),
          "max gap", d["pipeline"]["max_unit_gap_ms"], " ".join(f"{k[:6]}={v}" for k, v in st.items()),
          "sum", round(sum(v for v in st.values() if v > 0), 3), flush=True)
except Exception as e:
    print(name, "FAILED", e, open(f"gpurun_out/bench_{name}.err").read()[-800:])
PY
done
