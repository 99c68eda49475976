#!/bin/sh
# kernel tuning sweep on the bench workload (run on the GPU box)
for cfg in "1x8 0 16 32" "2x4 0 16 32" "4x2 0 16 32" "8x1 0 16 32" "1x8 1 16 32" "2x4 1 16 32" "4x2 1 16 32" "1x4 0 16 32" "1x8 0 32 16" "1x8 0 8 64" "2x4 0 32 16"; do
  set -- $cfg
  ALISB_BLOCK=$1 ALISB_STREAM=$2 ALISB_WAVES=$3 ALISB_MINCHUNK=$4 python bench.py --steps 10 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
l=json.loads(sys.stdin.readlines()[-1]); r=l['roofline']
print('cfg $cfg', 'value %.0f Mcells/s' % l['value'], 'ms/step %.3f' % l['ms_per_step'], 'interior GB/s %.0f frac %.3f' % (r['achieved'], r['frac']))"
done
