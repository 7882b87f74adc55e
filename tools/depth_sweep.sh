show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), d['run']['timed_batches'], d.get('corner_cut',{}).get('rows_skipped'))"; }
for dp in 1 2 3 4 5 6 8; do python bench.py --quick --steps 24 --warmup 4 --batches 3 --depth $dp 2>/dev/null | show depth$dp; done
