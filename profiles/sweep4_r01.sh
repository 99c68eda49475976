#!/bin/sh
# sub-launch concurrency sweep
for cfg in "0 256 4" "1 256 4" "2 256 4" "1 512 2" "2 512 2" "1 100000 8"; do
  set -- $cfg
  ALISB_CONC=$1 ALISB_WAVES=$2 ALISB_MINCHUNK=$3 python bench.py --steps 10 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readlines()[-1]); r=l['roofline']
print('conc/waves/minchunk $cfg', 'value %.0f Mcells/s' % l['value'], 'ms/step %.3f' % l['ms_per_step'], 'interior ms %.3f GB/s %.0f frac %.3f' % (r['avg_launch_ms'], r['achieved'], r['frac']))"
done
