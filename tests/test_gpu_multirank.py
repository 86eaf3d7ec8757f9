"""N-rank == 1-rank parity over NCCL (needs >= 2 GPUs; skipped on the single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["box", "rotbox-metis", "ogrid-metis", "warpbox-rcb"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_operator_matches_single_gpu(lib_built, world, case):
    """box: structured block split (fused path); rotbox-metis: METIS partition of rotated parallelepipeds (fused path,
    irregular halo); ogrid-metis: config C2's O-grid with boundary conditions (general path); warpbox-rcb: RCB."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(here, "multirank_worker.py"), case]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("case,method", [("quad-gll", "metis"), ("quad-gl3", "rcb"), ("axisym-argon6", "metis"), ("quad-nr", "metis")])
@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_generic_path_matches_single_gpu(lib_built, world, case, method):
    """The generic tensor-product path (2-D quadrilaterals / mixtures / axisymmetric) on irregular partitions."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29627", os.path.join(here, "multirank_generic_worker.py"), case, method]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
