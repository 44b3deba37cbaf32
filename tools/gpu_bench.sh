#!/bin/bash
# what the driver runs at round end, on one GPU: the GPU test-suite, smoke(), both bench arms
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -4 gpurun_out/r2_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py --impl reference ${BENCH_ARGS} ) > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -3 gpurun_out/r2_bench_ref.err
( time timeout 1500 python bench.py ${BENCH_ARGS} ) > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -5 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_ref.json", "gpurun_out/r2_bench_n1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, json.dumps({k: d[k] for k in d if k not in ("config",)}, indent=None)[:6000])
    except Exception as e:
        print(f, "unparsable", e)
PY
