for c in 8 4 12 16 8; do
  echo "== chunk $c"
  python bench.py --streams 48 --chunk $c --steps 6 --warmup 2 --no-e2e --no-cpu --no-summed 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('%.3f G  %.2f ms  sm %s chunk %s' % (d['value']/1e9, d['ms_per_step'], d['clocks']['sm_mhz'], d['config']['chunk_streams']))
"
done
