#!/bin/bash
# usage: gpurun --gpus N -- 'bash tools/gpu_multi.sh N'   : what the driver runs for the scaling step (both arms)
N=${1:-2}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --impl reference --gpus $N ${BENCH_ARGS} ) > gpurun_out/r2_bench_ref_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err
( time timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N ${BENCH_ARGS} ) > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -4 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
for f in ("gpurun_out/r2_bench_ref_n$N.json", "gpurun_out/r2_bench_n$N.json"):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        keep = {k: d[k] for k in ("value", "ms_per_step", "n_gpus") if k in d}
        keep["e2e"] = {k: v for k, v in d.get("e2e", {}).items() if not isinstance(v, str)}
        r = d.get("roofline") or {}
        keep["roofline"] = {k: r.get(k) for k in ("frac", "kernel_ms", "parity_max_err")}
        c3 = r.get("cfg3") or {}
        keep["cfg3"] = {k: c3.get(k) for k in ("value", "frac", "kernel_ms", "ms_per_step", "parity_max_err", "loglik", "loglik_rel_err_vs_n1", "loglik_evals_per_s")}
        print(f, json.dumps(keep))
    except Exception as e:
        print(f, "unparsable", e)
PY
