#!/bin/bash
# interleaved tile walk of neighbouring CTAs on the strided axes (MSM_B200_ILV) on the short 16-stream bench
for w in 1 2 4 8 1; do
  echo "== MSM_B200_ILV=$w"
  MSM_B200_ILV=$w python bench.py --streams 16 --steps 6 --warmup 2 --no-e2e --no-cpu --no-summed 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('%.3f G  %.2f ms  sm %s' % (d['value']/1e9, d['ms_per_step'], d['clocks']['sm_mhz']))
print('   '.join('%s=%.1f' % (k['name'].split('<512,')[1].rstrip('>'), k['ms']) for k in d['roofline']['kernels'][:11]))
"
done
