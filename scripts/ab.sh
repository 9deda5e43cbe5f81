# A/B of alternative builds of the library: SLAFB_LIB=<path> python bench.py ...
for lib in "$@"; do
  echo "== $lib"
  SLAFB_LIB=$lib timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; i=r['isolated']
        print('  ms/step %.4f value %.1fM | pipelined K2 %.4f K1 %.4f | isolated K2 %.4f (%.2f) K1 %.4f (%.2f)' % (d['ms_per_step'], d['value']/1e6, r['avg_launch_ms'], r['csr_build_kernel']['avg_launch_ms'], i['tokenize_cells_kernel']['avg_launch_ms'], i['tokenize_cells_kernel']['frac'], i['csr_build_kernel']['avg_launch_ms'], i['csr_build_kernel']['frac']))
    elif 'Error' in l or 'error' in l: print(l.rstrip())
"
done
