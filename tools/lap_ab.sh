show() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],4), 'single', d['single_pair_latency_ms'], {k:(v['ms'],v['gbs']) for k,v in d['stages'].items() if k.startswith('lap')})"; }
L=$PWD/karios_b200/_lib/libkarios_b200
python -m pytest tests -m gpu -x -q -k "laplacian or klt_match or float_raster or shape_and_dtype or full_s2_scene or auto_modes" 2>&1 | tail -1
for i in 1 2; do
for v in "$@"; do KR_LIB=${L}_$v.so python bench.py --quick --steps 24 --warmup 4 --batches 3 2>/dev/null | show $v; done
python bench.py --quick --steps 24 --warmup 4 --batches 3 2>/dev/null | show NEW
done
