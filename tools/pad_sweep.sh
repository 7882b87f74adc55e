show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), 'single', d['single_pair_latency_ms'], {k:v['ms'] for k,v in d['stages'].items() if k in ('corner_response','lk_roundtrip')}, d['pipeline'])"; }
for i in 1 2; do
for bps in 4 5; do KR_EIG_BPS=$bps python bench.py --quick --steps 24 --warmup 4 --batches 3 2>/dev/null | show bps$bps; done
for dp in 4 8; do python bench.py --quick --steps 24 --warmup 4 --batches 3 --depth $dp 2>/dev/null | show default_depth$dp; done
done
python -m pytest tests -m gpu -x -q -k "good_features or corner or running_cut or full_s2_scene" 2>&1 | tail -2
