# same-box A/B of library builds (KR_LIB) and of the running cut (KR_EIG_NOCUT)
show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), 'single', d['single_pair_latency_ms'], {k:v['ms'] for k,v in d['stages'].items() if k in ('corner_response','select')}, 'cand', d['stats_last']['n_candidates'], d.get('corner_cut',{}).get('rows_skipped'), d['pipeline'])"; }
L=$PWD/karios_b200/_lib/libkarios_b200
for i in 1 2; do
python bench.py --quick --steps 20 --warmup 3 --batches 3 2>/dev/null | show BASE
for v in "$@"; do KR_LIB=${L}_$v.so python bench.py --quick --steps 20 --warmup 3 --batches 3 2>/dev/null | show $v; done
done
