#!/bin/bash
# A/B of TPSB_COMM_GRAIN (persistent grids vs CTAs that retire after a few elements) in a partitioned run of N ranks.
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for g in ${GRAINS:-0 8}; do
  TPSB_COMM_GRAIN=$g run 2962$((g % 10)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ab_grain_${g}_n$N.json 2> gpurun_out/ab_grain_${g}_n$N.err
  python - gpurun_out/ab_grain_${g}_n$N.json $g <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k = d["roofline"].get("kernel_ms_per_step")
    print("grain", sys.argv[2], "%.4e" % d["value"], "%.3f ms" % d["ms_per_step"], "sum of timers %.3f" % sum(k.values()), k, d.get("multirank_parity") and d["multirank_parity"]["max_rel_diff"])
except Exception as ex:
    print("grain", sys.argv[2], "FAILED", ex)
PY
done
