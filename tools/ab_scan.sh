# A/B of the scan kernel generations on one GPU: device-resident steps only
for k in v3 pp; do
  for w in alt-grid null-grid perms; do
    BLMM_SCAN_KERNEL=$k python bench.py --workload $w --steps 10 --no-cpu --no-e2e --no-other 2> gpurun_out/ab_${k}_$w.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$k', '$w', 'ms/step %.3f' % d['ms_per_step'], 'kernel_ms %.3f' % d['roofline']['kernel_ms'], 'frac %.3f' % d['roofline']['frac'], d['clocks'])
"
  done
done
