for cfg in "16 8" "8 8" "4 4" "2 2"; do
  set -- $cfg
  echo "== streams $1 chunk $2"
  python bench.py --streams $1 --chunk $2 --steps 6 --warmup 2 --no-e2e --no-cpu --no-summed 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
ns=d['config']['streams_per_gpu']
print('%.3f G  %.2f ms/step  sm %s' % (d['value']/1e9, d['ms_per_step'], d['clocks']['sm_mhz']))
print('   '.join('%s=%.0f' % (k['name'].split('<512,')[1].rstrip('>'), k['GBps']) for k in d['roofline']['kernels'][:11]))
"
done
