# out-of-place launches under opaque region boxes: frame read skipped under opaque vectors or not
for c in 3 5; do for v in 0 1; do
  FLUC_TTMLBLEND_OPAQUE_SKIP=$v python bench.py --config $c --opaque-boxes --steps 200 --warmup 5 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg $c opaque boxes, OPAQUE_SKIP=$v:', round(d['value']), 'fps', round(d['ms_per_step'],5), 'ms/step, sustained', round(d['sustained']['value']))"
done; done
