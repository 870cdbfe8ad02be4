for f in RGBA BGRA AYUV; do
  for lib in build/libfluc_ttmlblend_before_packed.so flu-plugins-oss_b200/csrc/libfluc_ttmlblend.so; do
    FLUC_TTMLBLEND_LIB=$PWD/$lib python bench.py --config 4 --format $f --steps 200 --warmup 5 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$f', '$lib'.split('/')[-1], round(d['value']), 'fps', round(d['ms_per_step'],5), 'ms/step', round(d['roofline']['frac'],3), 'sust', round(d['sustained']['value']))"
  done
done
