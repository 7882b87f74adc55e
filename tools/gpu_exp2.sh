#!/bin/bash
# GPU experiment 2: full parity suite, tuning sweep, one ncu --set full capture of the big kernels.
mkdir -p gpurun_out
bash tools/gpu_check.sh nobench > gpurun_out/check.log 2>&1
grep -E "passed|failed|error" gpurun_out/pytest.log | sort | uniq -c | sort -rn | head -8
grep -B2 -A12 -E "FAILED|Error" gpurun_out/pytest.log | head -60
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
run lseg32 KR_LAP4_SEG=32
run lseg48 KR_LAP4_SEG=48
run lseg96 KR_LAP4_SEG=96
run eseg128 KR_EIG_SEG=128
run eseg192 KR_EIG_SEG=192
run eseg256 KR_EIG_SEG=256
run e3blk KR_EIG_SMEM_PAD=17408
run bps5 KR_EIG_BPS=5
run bps5s192 KR_EIG_BPS=5 KR_EIG_SEG=192
BENCH_ARGS="--depth 8" run d8bps5 KR_EIG_BPS=5
BENCH_ARGS="--depth 3" run depth3 X=1
BENCH_ARGS="--depth 2" run depth2 X=1
ncu --set full --clock-control none --import-source on \
    -k "regex:k_eig_approx|k_lk_roundtrip|k_laplacian4|k_pyr_down|k_exact_cands" -c 6 \
    -f -o gpurun_out/prof_v10 \
    python bench.py --steps 1 --warmup 0 --depth 1 --quick > gpurun_out/ncu_v10.log 2>&1
ls -la gpurun_out/prof_v10.ncu-rep
