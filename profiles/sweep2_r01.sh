#!/bin/sh
# chunking sweep (waves target / minimum planes per chunk) on the bench workload
ALISB_DEBUG=1 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | grep alisb | head -16
for cfg in "1 1024" "4 32" "8 32" "16 32" "32 32" "64 16" "128 8" "256 4"; do
  set -- $cfg
  ALISB_WAVES=$1 ALISB_MINCHUNK=$2 python bench.py --steps 10 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readlines()[-1]); r=l['roofline']
print('waves/minchunk $cfg', 'value %.0f Mcells/s' % l['value'], 'ms/step %.3f' % l['ms_per_step'], 'interior ms %.3f GB/s %.0f frac %.3f' % (r['avg_launch_ms'], r['achieved'], r['frac']))"
done
