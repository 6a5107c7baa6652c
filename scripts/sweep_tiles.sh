#!/bin/bash
# A/B of the tile-walk settings through their environment switches (no rebuild): bash scripts/sweep_tiles.sh
for cfg in "MSM_B200_TPCF=1" "MSM_B200_TPCF=2" "MSM_B200_TPCF=4" "MSM_B200_TPCF=8" "MSM_B200_TPCF=1" "MSM_B200_TPCF=4"; do
  echo "== $cfg"
  env $cfg python bench.py --streams 16 --steps 6 --warmup 2 --no-e2e --no-cpu --no-summed 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('%.3f G  %.2f ms  sm %s' % (d['value']/1e9, d['ms_per_step'], d['clocks']['sm_mhz']))
print('   '.join('%s=%.1f' % (k['name'].split('<512,')[1].rstrip('>'), k['ms']) for k in d['roofline']['kernels'][:11]))
"
done
