#!/bin/bash
# slow-axis block of the device layout (2^lb rows of i_lo per block): MSM_B200_LB sweep on the short 16-stream bench
for lb in 4 5 3 4; do
  echo "== MSM_B200_LB=$lb"
  MSM_B200_LB=$lb python bench.py --streams 16 --steps 6 --warmup 2 --no-e2e --no-cpu --no-summed 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('%.3f G  %.2f ms  sm %s' % (d['value']/1e9, d['ms_per_step'], d['clocks']['sm_mhz']))
print('   '.join('%s=%.1f' % (k['name'].split('<512,')[1].rstrip('>'), k['ms']) for k in d['roofline']['kernels'][:11]))
"
done
