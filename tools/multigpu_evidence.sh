#!/bin/bash
# Multi-GPU evidence for one world size N (run under `gpurun --gpus N`): N-rank == 1-rank parity of the four partition
# cases, the weak-scaled TGV bench line (with its checksum parity check) and the METIS-partitioned O-grid (config C2
# restated) bench line.  Writes gpurun_out/r2_multirank_nN.log, r2_bench_nN.json, r2_bench_cyl3d_nN.json.
N=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for c in box rotbox-metis ogrid-metis warpbox-rcb; do
  run 29617 tests/multirank_worker.py $c 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM"
  echo "exit ${PIPESTATUS[0]}"
done > gpurun_out/r2_multirank_n$N.log 2>&1
run 29618 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
run 29619 bench.py --gpus $N --steps 10 --warmup 3 --workload cyl3d --elems 48 > gpurun_out/r2_bench_cyl3d_n$N.json 2> gpurun_out/r2_bench_cyl3d_n$N.err
grep -h "rel_l2\|exit\|case" gpurun_out/r2_multirank_n$N.log | tail -30
for f in gpurun_out/r2_bench_n$N.json gpurun_out/r2_bench_cyl3d_n$N.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["n_gpus"], "%.4e" % d["value"], "%.3f ms" % d["ms_per_step"], "e2e %.3e" % (d["e2e"]["value"] or 0), d.get("multirank_parity") and d["multirank_parity"]["max_rel_diff"], d.get("partition"))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
