show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), 'single', d['single_pair_latency_ms'], {k:v['ms'] for k,v in d['stages'].items() if k in ('lk_roundtrip','zncc')})"; }
L=$PWD/karios_b200/_lib/libkarios_b200
for i in 1 2; do
python bench.py --quick --steps 24 --warmup 4 --batches 3 2>/dev/null | show BASE
for v in "$@"; do KR_LIB=${L}_$v.so python bench.py --quick --steps 24 --warmup 4 --batches 3 2>/dev/null | show $v; done
done
python -m pytest tests -m gpu -x -q -k "pyr_lk or klt_tracker or klt_match or zncc or full_s2_scene" 2>&1 | tail -1
