#!/usr/bin/env python
"""bench.py -- throughput of the DG right-hand-side path (RHSoperator::Mult) on B200.

Metric (BASELINE.json): RHS DOF-evals/s of the 3-D DG p=3 Navier-Stokes operator, DOF = one DG node.
A "step" is one RHSoperator::Mult over the whole mesh.  Workload: the synthetic periodic Taylor-Green
box of SURVEY.md section 8(d) (config C5): n^3 affine hexes per GPU (default n = 96 -> 56.6 M nodes,
283 M unknowns per GPU), weak-scaled over a 2x1x1 / 2x2x1 / 2x2x2 rank grid with NCCL face-neighbour
exchange.  Inputs (2.3 GB per field vector) are far larger than the 126 MB L2, so no L2 flush is needed
between iterations.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...     # the reference's CPU algorithm (oracle) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "rhs_dof_evals_per_s"
UNIT = "DOF-evals/s"
B_PER_DOF = 360.0  # 8*neq*(3+2*dim): U + write gradUp + U + gradUp + write dU/dt
# algorithmic bytes per DOF for each kernel class ("kernels and rooflines").  The library reports four timer
# classes; on the fused path (three launches, the default for p = 3 affine meshes) they are
#   grad       = elem_fused_kernel: U in 40, volume part of dU/dt out 40, face-trace blocks out 120
#   face_flux  = face_flux_mma_kernel: trace blocks in 120, face residuals out 30
#   elem_resid = lift_kernel: volume part in 40, face residuals in 30, dU/dt out 40
KERNEL_BYTES = {"fused": {"grad": 200.0, "face_flux": 150.0, "elem_resid": 110.0},
                "other": {"prim": 80.0, "grad": 160.0, "face_flux": 160.0, "elem_resid": 200.0}}
KERNEL_NAMES = {"fused": {"grad": "elem_fused_kernel", "face_flux": "face_flux_mma_kernel", "elem_resid": "lift_kernel"},
                "other": {"prim": "prim_kernel", "grad": "grad*_kernel", "face_flux": "face_flux*_kernel",
                          "elem_resid": "elem_resid_kernel"}}
# DRAM bytes per DOF (dram__bytes_read.sum + dram__bytes_write.sum) of one launch of each kernel, from the
# `ncu --set full` captures summarised in profiles/ (TGV 64^3, 16.8 M DG nodes)
NCU_DRAM_BYTES_PER_DOF = {"fused": {"grad": 228.9, "face_flux": 151.4, "elem_resid": 110.3},
                          "other": {"prim": 77.1, "grad": 347.5, "face_flux": 151.3, "elem_resid": 211.0}}
NCU_SOURCE = {"fused": "profiles/r2_ncu_full_summary.txt", "other": "profiles/r1q_ncu_full_summary.txt"}
PROC_GRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
PI = float(np.pi)


def tgv_visc_mult(Re=1600.0):
    """viscosityMultiplier giving Re = rho0 V0 L / mu = 1600 with L = 1 (box 2 pi), Sutherland mu at T0."""
    rho0, p0, gamma, R = 1.2, 101300.0, 1.4, 287.058
    T0 = p0 / (rho0 * R)
    V0 = 0.1 * np.sqrt(gamma * p0 / rho0)
    mu_s = 1.458e-6 * T0 ** 1.5 / (T0 + 110.4)
    return float(rho0 * V0 / Re / mu_s)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines, self.mark_at = device, None, [], 0

    def mark(self):
        """start of the timed region: nvidia-smi needs a few hundred ms to come up, so it is started before the warm-up and
        only the samples taken after this mark are reported"""
        self.mark_at = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = self.lines[self.mark_at:]
        in_region = len(window) > 0
        if not in_region:  # a timed region shorter than the sampling period: the warm-up samples (same load) stand in
            window = self.lines[-4:]
        for ln in window:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k] == "Active":
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": "timed region" if in_region else "warm-up (same load)"}


def cpu_baseline(n, nthreads, evals=3, eq=1, target_s=12.0):
    """The reference's CPU algorithm (dense per-element operators, per-quadrature-point physics calls)
    timed on the host cores: oracle/_ref (reference object code for the physics) when present, else the
    port.  Returns (DOF-evals/s, kind, ndofs)."""
    import oracle_api
    import tps_b200
    from common import tgv_state
    kind = "ref" if os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so")) else "port"
    if kind == "port" and not os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "liboracle.so")):
        oracle_api.build()
    m = tps_b200.cartesian_hex_mesh(n, n, n, lo=(-PI,) * 3, hi=(PI,) * 3)
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(eq, tgv_visc_mult()), nthreads=nthreads, kind=kind)
    U = tgv_state(orc.node_coords(), perturb=0.0)
    orc.mult(U)
    t0 = time.perf_counter()
    orc.mult(U)
    t1 = time.perf_counter() - t0
    evals = int(min(200, max(evals, target_s / max(t1, 1e-3))))  # ~target_s seconds of CPU work
    t0 = time.perf_counter()
    for _ in range(evals):
        orc.mult(U)
    dt = (time.perf_counter() - t0) / evals
    return orc.N / dt, ("reference" if kind == "ref" else "port"), orc.N, dt, evals


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores."""
    if rank != 0:
        return
    import oracle_api
    import tps_b200
    from common import tgv_state
    nthreads = os.cpu_count()
    n = args.cpu_n
    kind = "ref" if os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so")) else "port"
    if kind == "port" and not os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "liboracle.so")):
        oracle_api.build()
    m = tps_b200.cartesian_hex_mesh(n, n, n, lo=(-PI,) * 3, hi=(PI,) * 3)
    orc = oracle_api.Oracle(3, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(1, tgv_visc_mult()), nthreads=nthreads, kind=kind)
    U = tgv_state(orc.node_coords(), perturb=0.0)
    for _ in range(args.warmup):
        orc.mult(U)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.mult(U)
    el = time.perf_counter() - t0
    val = orc.N * args.steps / el
    sample = f"TGV box {n}^3 hexes p=3 ({orc.N} DG nodes), {args.steps} RHS evaluations, {nthreads} OpenMP threads"
    k = "reference" if kind == "ref" else "port"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, (1, 1, 1), sample_n=n),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": nthreads, "kind": k, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, grid, sample_n=None):
    n = args.n
    if getattr(args, "workload", "tgv") == "cyl3d":
        return {"workload": "C2 cyl3d restated on a hex O-grid (trilinear elements, inlet/outlet/isothermal wall), "
                            "DG p=3 GL/GL, dry-air Navier-Stokes", "global_elements": f"{n}x{4 * n}x{n}", "order": 3,
                "num_equation": 5, "rank_grid": "METIS k-way element partition" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else "1x1x1",
                "l2_policy": "inputs exceed the 126 MB L2 for n >= 32; no flush needed"}
    cfg = {"workload": "C5 synthetic periodic 3-D hex box (compressible Taylor-Green), DG p=3 GL/GL, dry-air "
                       "Navier-Stokes, Re=1600, M0=0.1",
           "elements_per_gpu": f"{n}^3", "global_elements": f"{n * grid[0]}x{n * grid[1]}x{n * grid[2]}",
           "order": 3, "num_equation": 5, "rank_grid": "x".join(map(str, grid)),
           "l2_policy": "inputs (2.3 GB per state vector at 96^3) exceed the 126 MB L2; no flush needed"}
    if getattr(args, "strong", 0):
        g = args.strong
        cfg["elements_per_gpu"] = f"{g // grid[0]}x{g // grid[1]}x{g // grid[2]}"
        cfg["global_elements"] = f"{g}x{g}x{g}"
    if sample_n is not None:
        cfg["cpu_sample_elements"] = f"{sample_n}^3"
    return cfg


# BASELINE configs C1 / C3 / C4 at the reference's sizes on the generic tensor-product path (algorithmic
# bytes per DOF-eval 64 / 336 / 616).  Single GPU; same JSON contract as the headline line.
GENERIC_WORKLOADS = {
    "c1": ("C1 mms.euler_2d: 2-D Euler, 160 x 160 periodic quadrilaterals (25 600), p = 2, Gauss-Lobatto nodes and rule", 64.0),
    "c3": ("C3 mms.ternary_2d: 2-D ternary argon plasma (Ar+, e, Ar; two temperatures), collision-integral transport and "
           "ionisation chemistry, 64 x 64 periodic quadrilaterals, p = 2, Gauss-Lobatto nodes and rule (the ini's discretisation; "
           "its constant-transport variant is covered by tests/test_gpu_plasma.py)", 336.0),
    "c4": ("C4 plasma.axisym type: axisymmetric six-species two-temperature argon, constant transport, inlet / outlet / "
           "inviscid + isothermal walls, 40 x 40 quadrilaterals, p = 3", 616.0),
}


def run_generic_workload(args):
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import tps_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the RHS path has no CPU fallback")
    torch.cuda.set_device(0)
    sampler = ClockSampler(0)
    sampler.start()
    t0 = time.perf_counter()
    PI = np.pi
    wl = args.workload
    orc = None
    if wl == "c1":
        n = args.n if args.n != 96 else 160
        m = tps_b200.cartesian_quad_mesh(n, n, lo=(0, 0), hi=(3.02, 3.02))
        op = tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.dry_air(0), basis_type=1, int_rule_type=1)
        from test_gpu_generic_parity import _state2d
        from common import node_coords_from_mesh  # noqa: F401
        import oracle_api
        orc = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                                phys=oracle_api.dry_air_params(0), basis_type=1, int_rule=1)
        U = _state2d(orc.node_coords() * (2 * PI / 3.02))
        cpu_kind = "port"
    else:
        import axisym_cases as ac
        import oracle_api
        import plasma_cases as pc
        if wl == "c3":
            n = args.n if args.n != 96 else 64
            m = tps_b200.cartesian_quad_mesh(n, n, lo=(-PI, -PI), hi=(PI, PI))
            pm = tps_b200.PlasmaModels.from_dict(pc.argon_minimal_dict())
            op = tps_b200.RhsOperator(m, order=2, physics=tps_b200.Physics.plasma_mixture(pm, 1), nvel=2, basis_type=1, int_rule_type=1)
            orc = oracle_api.Oracle(2, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                                    phys=oracle_api.mixture_params(pm, 1), kind="ref", neq=op.neq, nvel=2, basis_type=1, int_rule=1)
            up = pc.hot_primitives(orc.node_coords())
        else:
            n = args.n if args.n != 96 else 40
            m = ac.box(n=(n, n), warp=0.05)
            op, orc = ac.make_pair(m, 3, 1, 0, 0, 3, "c4", True, mixture=ac.argon6_dict())
            up = ac.argon6_primitives(orc.node_coords(), 3)
        U = np.ascontiguousarray(orc.pt("cons", up).T).reshape(-1)
        cpu_kind = "reference"
    setup_s = time.perf_counter() - t0
    N, neq = op.N, op.neq
    x = torch.from_numpy(U).cuda()
    y = torch.empty_like(x)
    steps, warmup = args.steps, max(args.warmup, 3)
    flush = torch.empty(160 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda")  # > 126 MB L2: these states are tiny
    for _ in range(warmup):
        op.Mult(x, y)
    torch.cuda.synchronize()
    sampler.mark()
    l0 = op.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        flush.fill_(1.0)
        e0.record()
        op.Mult(x, y)
        e1.record()
    torch.cuda.synchronize()
    launches = op.launch_count() - l0
    clocks = sampler.stop()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in evs) / steps
    value = N / (ms * 1e-3)
    op.set_profiling(True)
    op.kernel_times()
    for _ in range(3):
        op.Mult(x, y)
    kt = {k: v[0] / 3 for k, v in op.kernel_times().items() if v[1] > 0}
    op.set_profiling(False)
    # parity of this very evaluation against the oracle (the sizes are small enough)
    yo = orc.mult(U)
    yd = y.cpu().numpy()
    err = max(float(np.linalg.norm(yd[k * N:(k + 1) * N] - yo[k * N:(k + 1) * N]) / max(np.linalg.norm(yo[k * N:(k + 1) * N]), 1e-300))
              for k in range(neq))
    # CPU beside it: the oracle on the same mesh (reference physics object code for the plasma cases), all host threads
    tc = time.perf_counter()
    reps = 0
    while time.perf_counter() - tc < 10.0 and reps < 50:
        orc.mult(U)
        reps += 1
    cpu_s = (time.perf_counter() - tc) / reps
    # end to end: host buffers through tpsb_rhs_mult_host
    hx, hy = torch.from_numpy(U).pin_memory(), torch.empty(U.shape, dtype=torch.float64).pin_memory()
    for _ in range(2):
        op.mult_host(hx, hy)
    te = time.perf_counter()
    e2e_steps = 10
    for _ in range(e2e_steps):
        op.mult_host(hx, hy)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - te) / e2e_steps * 1e3
    peak, peak_src = peaks()
    desc, bytes_per_dof = GENERIC_WORKLOADS[wl]
    dom = max(kt, key=kt.get) if kt else "elem_resid"
    out = {"metric": "rhs_dof_evals_per_s", "value": value, "unit": "DOF-evals/s", "n_gpus": 1, "steps": steps, "warmup": warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": desc, "elements": int(op.NE), "order": int(3 if wl == "c4" else 2), "num_equation": int(neq),
                      "l2_policy": "160 MB buffer written between timed evaluations (the state fits in L2)"},
           "roofline": {"bound": "hbm", "kernel": "gen_resid_kernel" if dom == "elem_resid" else "gen_grad_kernel", "timer_class": dom,
                        "achieved": bytes_per_dof * value / 1e9, "peak": peak, "unit": "GB/s", "frac": bytes_per_dof * value / 1e9 / peak,
                        "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_dof": bytes_per_dof,
                        "kernel_ms_per_step": kt,
                        "note": "whole-step figure on BASELINE.md section 3's bytes per DOF-eval; these configurations are bound by "
                                "per-point physics and dependent latency at one CTA per element, not by HBM"},
           "cpu_baseline": {"value": N / cpu_s, "unit": "DOF-evals/s", "cores": os.cpu_count(), "kind": cpu_kind,
                            "sample": f"the same mesh and state, {reps} evaluations of {cpu_s:.3f} s (oracle with "
                                      + ("the reference's physics object code" if cpu_kind == "reference" else "the restated dry-air physics") + ")"},
           "clocks": clocks,
           "e2e": {"value": N / (e2e_ms * 1e-3), "unit": "DOF-evals/s", "h2d_bytes_per_step": int(U.nbytes), "d2h_bytes_per_step": int(U.nbytes),
                   "steps": e2e_steps, "api": "tpsb_rhs_mult_host (pinned host x -> device, Mult, y -> pinned host)"},
           "gpu_launches": int(launches), "finite": bool(np.isfinite(yd).all()), "setup_s": setup_s, "dofs_per_gpu": int(N),
           "path": op.path(), "parity_vs_oracle_rel_l2": err}
    if not (err < 1e-9):
        raise SystemExit(f"--workload {wl}: device result differs from the oracle ({err:.3e})")
    os.dup2(saved_stdout, 1)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--elems", dest="n", type=int, default=96,
                    help="elements per direction per GPU (--elems: the spelling torchrun does not mistake for its own --n* flags)")
    ap.add_argument("--cpu-n", type=int, default=20, help="elements per direction of the CPU-baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--strong", type=int, default=0, metavar="G",
                    help="strong scaling: a fixed global G^3 box split over the ranks (SURVEY.md 8d: 128); "
                         "default 0 = weak scaling with --n elements per direction per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--partitioner", default="metis", choices=["metis", "rcb"], help="element partition of --workload cyl3d at N > 1")
    ap.add_argument("--no-verify", action="store_true", help="skip the N-rank == 1-rank checksum comparison at N > 1")
    ap.add_argument("--workload", default="tgv", choices=["tgv", "cyl3d", "c1", "c3", "c4"],
                    help="tgv: BASELINE config C5 (the headline line); cyl3d: config C2 restated on a hex O-grid "
                         "(general trilinear path + boundary conditions), single GPU, development measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.workload in ("c1", "c3", "c4"):
        if world > 1:
            raise SystemExit("--workload c1 / c3 / c4 are single-GPU lines (multi-rank parity of the generic path: tests/multirank_generic_worker.py)")
        run_generic_workload(args)
        return

    # the JSON line must be the only thing on stdout: libraries (NCCL prints its version banner there) write to
    # stderr until the final print
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import tps_b200
    from tps_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the RHS path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank)  # started now: nvidia-smi takes a few hundred ms to deliver its first sample
    if rank == 0:
        sampler.start()
    if world not in PROC_GRID:
        raise SystemExit(f"unsupported world size {world}")
    grid = PROC_GRID[world]
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            buf = capi.C.create_string_buffer(128)
            assert tps_b200.lib().tpsb_comm_get_unique_id(buf) == 0
            uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        cptr = capi.C.c_void_p()
        rc = tps_b200.lib().tpsb_comm_init_rank(bytes(uid.cpu().numpy().tobytes()), world, rank, local_rank,
                                                capi.C.byref(cptr))
        assert rc == 0, "tpsb_comm_init_rank failed"
        comm = cptr

    n = args.n
    gn = (n * grid[0], n * grid[1], n * grid[2])
    lo = tuple(-PI * g for g in grid)
    hi = tuple(PI * g for g in grid)
    if args.strong:  # fixed global box: the per-rank block shrinks with the rank grid
        assert all(args.strong % g == 0 for g in grid), "global size must divide by the rank grid"
        gn = (args.strong,) * 3
        lo, hi = (-PI,) * 3, (PI,) * 3
    phys = tps_b200.Physics.dry_air(1, tgv_visc_mult())
    t_setup = time.perf_counter()
    if args.workload == "cyl3d":
        # config C2 restated on a hex O-grid; N > 1: STRONG scaling of the one mesh under a METIS k-way element partition
        # (what MFEM's GeneratePartitioning gives M2ulPhyS, src/M2ulPhyS.cpp:332), irregular halos
        mesh = tps_b200.cylinder_ogrid_mesh(n, 4 * n, n, order_mode=1)
        specs = [(1, 2, 3, (300.0,)), (2, 0, 2, (1.2, 20.0, 0.0, 0.0)), (3, 1, 0, (101300.0,))]
        kw = dict(order=3, physics=tps_b200.Physics.dry_air(1, 50.0), device=local_rank, use_bc_in_grad=True,
                  bcs=[tps_b200.BcDesc.make(*b) for b in specs])
        if world == 1:
            op = tps_b200.RhsOperator(mesh, face_attr=mesh["face_attr"], **kw)
            NE = 4 * n ** 3
        else:
            elem_rank, cut = tps_b200.partition_elements(mesh, world, args.partitioner)
            mesh = tps_b200.partition_mesh(mesh, elem_rank, rank)
            halo = tps_b200.make_halo_desc(mesh, comm)
            op = tps_b200.RhsOperator(mesh, face_attr=mesh["face_attr"], halo=halo, num_nbr_elems=mesh["num_nbr_elems"], **kw)
            NE = mesh["num_elems"]
            args.partition_info = {"method": args.partitioner, "edge_cut": cut, "elements": np.bincount(elem_rank).tolist()}
    elif world == 1:
        mesh = tps_b200.cartesian_hex_mesh(gn[0], gn[1], gn[2], lo=lo, hi=hi, order_mode=1)
        op = tps_b200.RhsOperator(mesh, order=3, physics=phys, device=local_rank)
        NE = gn[0] * gn[1] * gn[2]
    else:
        mesh = tps_b200.cartesian_hex_partition(gn, grid, rank, lo=lo, hi=hi, order_mode=1)
        halo = tps_b200.make_halo_desc(mesh, comm)
        op = tps_b200.RhsOperator(mesh, order=3, physics=phys, device=local_rank, halo=halo,
                                  num_nbr_elems=mesh["num_nbr_elems"])
        NE = mesh["num_elems"]
    t_setup = time.perf_counter() - t_setup
    N = op.N

    # Taylor-Green initial state evaluated on the device from the element vertices
    T = capi.ref_tables(3)
    xn = torch.from_numpy(T["xn"]).to(dev)
    ii = torch.arange(64, device=dev)
    xi = torch.stack([xn[ii % 4], xn[(ii // 4) % 4], xn[ii // 16]], 1)
    hv = torch.tensor([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]],
                      dtype=torch.float64, device=dev)
    shp = torch.ones(64, 8, dtype=torch.float64, device=dev)
    for d in range(3):
        shp = shp * torch.where(hv[None, :, d] > 0, xi[:, None, d], 1.0 - xi[:, None, d])
    exyz = mesh["elem_xyz"][:NE]
    if world > 1 and args.workload == "tgv" and not args.strong:
        # Every block of the weak-scaled box sees the same 2 pi-periodic fields.  Evaluate them at the coordinates the
        # N = 1 box gives the element (same expression as meshkit: lo + index * h), so that every rank's input is
        # bit-identical to the single-GPU input and the checksum comparison below tests the operator, not sin(x + 2 pi).
        gid = mesh["elem_gid"][:NE]
        gi = np.stack([gid % gn[0], (gid // gn[0]) % gn[1], gid // (gn[0] * gn[1])], 1) % n          # [NE, 3] block-local
        hvn = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
        h1 = (PI - (-PI)) / n
        exyz = -PI + (gi[:, None, :] + hvn[None, :, :]).astype(np.float64) * h1
    ev = torch.from_numpy(np.ascontiguousarray(exyz)).to(dev)
    U = torch.empty(5 * N, dtype=torch.float64, device=dev)
    rho0, p0, gamma = 1.2, 101300.0, 1.4
    V0 = 0.1 * float(np.sqrt(gamma * p0 / rho0))
    chunk = 1 << 17
    for e0 in range(0, NE, chunk):
        e1 = min(NE, e0 + chunk)
        X = torch.einsum("na,ead->end", shp, ev[e0:e1]).reshape(-1, 3)
        x, y, z = X[:, 0], X[:, 1], X[:, 2]
        if args.workload == "cyl3d":  # potential-flow-like start around the cylinder (u -> 20 m/s far away)
            r2 = x * x + y * y
            u = 20.0 * (1.0 - 0.25 / r2)
            v = 2.0 * torch.sin(3 * x) * torch.cos(2 * y) * (1.0 - 0.25 / r2)
            p = 102300.0 + 50.0 * torch.cos(x) * torch.cos(y)
        else:
            u = V0 * torch.sin(x) * torch.cos(y) * torch.cos(z)
            v = -V0 * torch.cos(x) * torch.sin(y) * torch.cos(z)
            p = p0 + rho0 * V0 * V0 / 16.0 * (torch.cos(2 * x) + torch.cos(2 * y)) * (torch.cos(2 * z) + 2.0)
        sl = slice(e0 * 64, e1 * 64)
        U[0 * N:1 * N][sl] = rho0
        U[1 * N:2 * N][sl] = rho0 * u
        U[2 * N:3 * N][sl] = rho0 * v
        U[3 * N:4 * N][sl] = 0.0
        U[4 * N:5 * N][sl] = p / (gamma - 1.0) + 0.5 * rho0 * (u * u + v * v)
    del ev
    Y = torch.empty_like(U)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        op.Mult(U, Y)
    barrier()
    l0 = op.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    ev0.record()
    for _ in range(args.steps):
        op.Mult(U, Y)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = op.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    clocks = sampler.stop() if rank == 0 else None
    finite = bool(torch.isfinite(Y).all().item())
    N_global = N * world
    if world > 1:
        ng = torch.tensor([N], dtype=torch.int64, device=dev)
        dist.all_reduce(ng, op=dist.ReduceOp.SUM)
        N_global = int(ng.item())
    value = N_global * args.steps / (ms * 1e-3)

    # Order-independent checksums of dU/dt per equation: sum |y| and sum y^2.  In the weak-scaled periodic Taylor-Green
    # box every rank's block is identical to the N = 1 box (the fields have period 2 pi, the blocks are 2 pi wide), so
    # at N > 1 each rank must reproduce the single-GPU checksums: rank 0 evaluates the N = 1 operator on the same block
    # and the run FAILS if any rank differs by more than 1e-12 -- the scaling line carries N-rank == 1-rank parity.
    def checksums(y, n_local):
        yy = y.view(5, n_local)
        return torch.cat([yy.abs().sum(1), (yy * yy).sum(1)])
    cs = checksums(Y, N)
    cs_global = cs.clone()
    if world > 1:  # sums over all ranks: element-order and partition independent, comparable across N for a fixed mesh
        dist.all_reduce(cs_global, op=dist.ReduceOp.SUM)
    parity = None
    if world > 1 and args.workload == "tgv" and not args.strong and not args.no_verify:
        allcs = [torch.empty_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        ref = torch.empty_like(cs)
        if rank == 0:
            m1 = tps_b200.cartesian_hex_mesh(n, n, n, lo=(-PI,) * 3, hi=(PI,) * 3, order_mode=1)
            op1 = tps_b200.RhsOperator(m1, order=3, physics=phys, device=local_rank)
            ev1 = torch.from_numpy(np.ascontiguousarray(m1["elem_xyz"])).to(dev)
            U1 = torch.empty_like(U)
            for e0 in range(0, n ** 3, chunk):
                e1 = min(n ** 3, e0 + chunk)
                X = torch.einsum("na,ead->end", shp, ev1[e0:e1]).reshape(-1, 3)
                x, y, z = X[:, 0], X[:, 1], X[:, 2]
                u = V0 * torch.sin(x) * torch.cos(y) * torch.cos(z)
                v = -V0 * torch.cos(x) * torch.sin(y) * torch.cos(z)
                p = p0 + rho0 * V0 * V0 / 16.0 * (torch.cos(2 * x) + torch.cos(2 * y)) * (torch.cos(2 * z) + 2.0)
                sl = slice(e0 * 64, e1 * 64)
                U1[0 * N:1 * N][sl] = rho0
                U1[1 * N:2 * N][sl] = rho0 * u
                U1[2 * N:3 * N][sl] = rho0 * v
                U1[3 * N:4 * N][sl] = 0.0
                U1[4 * N:5 * N][sl] = p / (gamma - 1.0) + 0.5 * rho0 * (u * u + v * v)
            Y1 = op1.Mult(U1)
            ref = checksums(Y1, N)
            op1.close()
            del ev1, U1, Y1
        dist.broadcast(ref, 0)
        # d(rho)/dt of the initially solenoidal Taylor-Green field is a ~1e-5 cancellation residue of its flux divergence
        # (sum|d rho/dt| ~ 1e3 against sum|d(rho u)/dt| / V0 ~ 1e8), so its checksums are compared on the scale of that
        # flux divergence; every other equation on its own scale
        scale = ref.clone()
        scale[0], scale[5] = ref[1] / V0, ref[6] / (V0 * V0)
        per_rank = [((c - ref).abs() / scale) for c in allcs]
        worst = max(float(d.max().item()) for d in per_rank)
        raw = max(float(((c - ref).abs() / ref).max().item()) for c in allcs)
        parity = {"criterion": "per-equation sum|dU/dt| and sum (dU/dt)^2 of every rank vs the single-GPU operator on the "
                               "same block (bit-identical inputs); d(rho)/dt on the scale sum|d(rho u)/dt| / V0 of its flux "
                               "divergence", "max_rel_diff": worst, "max_rel_diff_unscaled": raw, "tol": 1e-12,
                  "ok": worst <= 1e-12, "per_equation_rel_diff": [float(t) for t in torch.stack(per_rank).max(0).values.tolist()],
                  "reference_checksum": [float(t) for t in ref.tolist()]}
        if not parity["ok"]:
            raise SystemExit(f"multi-rank parity FAILED: max relative checksum difference {worst:.3e} > 1e-12")

    # per-kernel device timers (separate, untimed pass): the dominant kernel's own roofline
    op.set_profiling(True)
    op.kernel_times()
    for _ in range(3):
        op.Mult(U, Y)
    kt = op.kernel_times()
    op.set_profiling(False)
    per_launch = {k: (v[0] / max(v[1], 1)) for k, v in kt.items() if v[1] > 0}
    per_step = {k: v[0] / 3.0 for k, v in kt.items() if v[1] > 0}
    kind = "fused" if op.path() == "fused" else "other"
    KB, NCU = KERNEL_BYTES[kind], NCU_DRAM_BYTES_PER_DOF[kind]
    dom = max((k for k in per_step if k in KB), key=lambda k: per_step[k])
    peak, peak_src = peaks()
    dom_launches_per_step = kt[dom][1] / 3.0
    dom_bytes_per_launch = KB[dom] * N / dom_launches_per_step
    achieved = dom_bytes_per_launch / (per_launch[dom] * 1e-3) / 1e9
    traffic = None
    if args.workload == "tgv" and dom in NCU:
        traffic = NCU[dom] * N / dom_launches_per_step  # bytes per launch, like `achieved`
    roofline = {"bound": "hbm", "kernel": KERNEL_NAMES[kind][dom], "timer_class": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": NCU_SOURCE[kind] + " (ncu --set full, bytes/DOF x DOFs of this launch)",
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes_per_launch, "launch_ms": per_launch[dom],
                "kernel_share_of_step": per_step[dom] / sum(per_step.values()),
                "kernel_ms_per_step": per_step}
    step_gbs = B_PER_DOF * (N * args.steps / (ms * 1e-3)) / 1e9  # per GPU
    roofline_step = {"bound": "hbm", "bytes_per_dof_eval": B_PER_DOF, "achieved": step_gbs, "peak": peak,
                     "unit": "GB/s", "frac": step_gbs / peak, "frac_of_nominal_8TBs": step_gbs / 8000.0}

    # end to end through the C ABI with HOST buffers (pinned), copies inside the timed region
    e2e = None
    try:
        hx = torch.empty(5 * N, dtype=torch.float64, pin_memory=True)
        hy = torch.empty(5 * N, dtype=torch.float64, pin_memory=True)
        hx.copy_(U)
        for _ in range(3):  # warm-up: pinned pages touched, copy streams and the chunk schedule's events created
            op.mult_host(hx, hy)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            op.mult_host(hx, hy)
        barrier()
        el = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([el], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            el = float(t.item())
        e2e = {"value": N_global * args.e2e_steps / el, "unit": UNIT, "h2d_bytes_per_step": int(5 * N * 8) * world,
               "d2h_bytes_per_step": int(5 * N * 8) * world, "steps": args.e2e_steps,
               "api": "tpsb_rhs_mult_host (pinned host x -> device, Mult, y -> pinned host)"}
        del hx, hy
    except Exception as ex:  # pinned allocation can fail on small hosts
        e2e = {"value": None, "unit": UNIT, "error": str(ex)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nthreads = os.cpu_count()
        val, kind, nd, dt, nev = cpu_baseline(args.cpu_n, nthreads)
        cpu = {"value": val, "unit": UNIT, "cores": nthreads, "kind": kind,
               "sample": f"TGV box {args.cpu_n}^3 hexes p=3 ({nd} DG nodes), {nev} RHS evaluations of {dt:.2f} s, "
                         f"{nthreads} OpenMP threads; dense per-element operators as in the reference CPU path"}

    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (args.strong or (args.workload == "cyl3d" and world > 1)) else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, grid),
            "roofline": roofline, "roofline_step": roofline_step, "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e,
            "gpu_launches": launches, "finite": finite, "setup_s": t_setup, "dofs_per_gpu": N, "path": op.path(),
            "checksum": [float(t) for t in cs.tolist()], "checksum_all_ranks": [float(t) for t in cs_global.tolist()],
            "multirank_parity": parity,
            "partition": getattr(args, "partition_info", None),
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
