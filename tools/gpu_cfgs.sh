#!/bin/bash
# bench lines of the other BASELINE configs at N = 1 (full sizes)
mkdir -p gpurun_out
for c in cfg1 cfg4 cfg5; do
  timeout 900 python bench.py --workload $c --steps 10 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
  tail -1 gpurun_out/r2_bench_$c.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$c', 'value %.4g' % d['value'], 'ms/step %.4g' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['roofline']['kernel'], 'frac %.3f' % d['roofline']['frac'],
      'parity %.2g' % d['roofline']['parity_max_err'], 'cpu %.4g' % d['cpu_baseline']['value'], 'loglik/s %.4g' % d['e2e']['loglik_evals_per_s'])"
done
