for te in 1 4 1000; do for foc in 0 1; do
echo "timing_every=$te front_on_caller=$foc"; SLAFB_BENCH_TIMING_EVERY=$te SLAFB_FRONT_ON_CALLER=$foc timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('  ms/step %.4f value %.1fM  K2 %.4f ms K1 %.4f ms gen %.4f host %s' % (d['ms_per_step'], d['value']/1e6, r['avg_launch_ms'], r['csr_build_kernel']['avg_launch_ms'], r['general_path']['avg_launch_ms'], d['host_us_per_step']))
"
done; done
