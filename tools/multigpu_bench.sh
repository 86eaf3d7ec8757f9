#!/bin/bash
# Bench lines only for one world size N (run under `gpurun --gpus N`): the weak-scaled TGV line (with its N-rank == 1-rank
# checksum check) and the METIS-partitioned O-grid.  Writes gpurun_out/r2_bench_nN.json, r2_bench_cyl3d_nN.json.
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29618 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
FILES=gpurun_out/r2_bench_n$N.json
if [ "$2" != "tgv" ]; then
  run 29619 bench.py --gpus $N --steps 10 --warmup 3 --workload cyl3d --elems 48 > gpurun_out/r2_bench_cyl3d_n$N.json 2> gpurun_out/r2_bench_cyl3d_n$N.err
  FILES="$FILES gpurun_out/r2_bench_cyl3d_n$N.json"
fi
for f in $FILES; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], "%.4e" % d["value"], "%.3f ms" % d["ms_per_step"], "e2e %.3e" % (d["e2e"]["value"] or 0), d.get("multirank_parity") and d["multirank_parity"]["max_rel_diff"], d["roofline"].get("kernel_ms_per_step"), d["clocks"])
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
