#!/bin/sh
# finer chunking sweep + block shape at small chunks
for cfg in "256 4 1x8" "512 2 1x8" "1024 1 1x8" "100000 1 1x8" "512 2 2x4" "512 2 1x4" "256 4 2x4" "100000 8 1x8" "100000 4 1x8" "100000 2 1x8"; do
  set -- $cfg
  ALISB_WAVES=$1 ALISB_MINCHUNK=$2 ALISB_BLOCK=$3 python bench.py --steps 10 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readlines()[-1]); r=l['roofline']
print('waves/minchunk/block $cfg', 'value %.0f Mcells/s' % l['value'], 'ms/step %.3f' % l['ms_per_step'], 'interior ms %.3f GB/s %.0f frac %.3f' % (r['avg_launch_ms'], r['achieved'], r['frac']))"
done
