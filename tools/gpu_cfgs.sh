#!/bin/bash
# bench lines of the other BASELINE configs at N = 1 (full sizes)
mkdir -p gpurun_out
for c in cfg1 cfg3 cfg4 cfg5; do
  timeout 900 python bench.py --workload $c --steps 10 --warmup 3 > gpurun_out/bench_$c.log 2> gpurun_out/bench_$c.err
  tail -1 gpurun_out/bench_$c.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$c', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['cpu_baseline']['value'], d['extras'].get('e2e_csc_sets_per_s'), d['extras'].get('loglik_evals_per_s'))"
done
