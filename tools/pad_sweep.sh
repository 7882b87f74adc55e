show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), 'single', d['single_pair_latency_ms'], {k:v['ms'] for k,v in d['stages'].items() if k in ('corner_response','lk_roundtrip')}, d['pipeline'])"; }
for pad in 0 2048 4096 8192; do KR_EIG_SMEM_PAD=$pad python bench.py --quick --steps 24 --warmup 4 --batches 3 2>/dev/null | show eigpad$pad; done
