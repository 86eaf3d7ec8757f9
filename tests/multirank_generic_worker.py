"""Worker for the generic-path multi-rank parity test (and manual `gpurun --gpus N` runs): the generic tensor-product
kernels (2-D quadrilaterals, Gauss-Lobatto, boundary conditions, axisymmetric six-species argon with the mixing-length
model) on an irregular METIS / RCB partition with NCCL face-neighbour exchange, against the same operator on the whole
mesh on one device.  Criterion of the reference's multi-rank regression tests: N-rank result == 1-rank result."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    import torch
    import torch.distributed as dist

    import axisym_cases as ac
    import tps_b200
    from common import rel_l2
    from tps_b200 import capi

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf = capi.C.create_string_buffer(128)
        assert tps_b200.lib().tpsb_comm_get_unique_id(buf) == 0
        uid.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    comm = capi.C.c_void_p()
    assert tps_b200.lib().tpsb_comm_init_rank(bytes(uid.cpu().numpy().tobytes()), world, rank, local_rank,
                                              capi.C.byref(comm)) == 0
    case = sys.argv[1] if len(sys.argv) > 1 else "quad-gll"
    method = sys.argv[2] if len(sys.argv) > 2 else "metis"
    # (mesh, order, basis, rule, nvel, bc set, useBCinGrad, mixture, mixing length)
    if case == "quad-gll":      # config C1 type + walls / inlet / outlet: p = 2, Gauss-Lobatto basis and rule, Navier-Stokes
        cfg = (ac.box(n=(9, 7), warp=0.05), 2, 1, 1, 2, "c4", True, None, None)
    elif case == "quad-gl3":    # Gauss-Legendre p = 3 quadrilaterals, all-wall box
        cfg = (ac.box(n=(8, 6), warp=0.04), 3, 0, 0, 2, "adiabatic", False, None, None)
    elif case == "quad-nr":     # non-reflecting inlet + mass-flow outlet: patch means and areas all-reduced over the ranks
        cfg = (ac.box(n=(9, 7), warp=0.05), 3, 0, 0, 2, "nr", True, None, None)
    elif case == "axisym-argon6":  # config C4 type: axisymmetric, six species, two temperatures, mixing length
        cfg = (ac.box(n=(7, 6), warp=0.03), 2, 1, 1, 3, "c4", True, ac.argon6_dict(), (0.05, 0.9, 0.3))
    else:
        raise SystemExit(f"unknown case {case}")
    gm, order, bt, ir, nvel, bck, ubg, mixture, ml = cfg
    elem_rank, cut = tps_b200.partition_elements(gm, world, method)
    part = tps_b200.partition_mesh(gm, elem_rank, rank)
    gop, _ = ac.make_pair(gm, order, 1, bt, ir, nvel, bck, ubg, mixture=mixture, mixing_length=ml, want_oracle=False, device=local_rank)
    # the global state: dry air from the smooth field, mixtures from primitives through the device's own conversion
    # nodal coordinates of the quadrilaterals from the 1-D nodes of the basis
    T = capi.ref_tables(order)
    if bt == 1:  # Gauss-Lobatto nodes on [0, 1]: end points and the roots of P'_p
        from numpy.polynomial import legendre
        inner = np.sort(legendre.Legendre.basis(order).deriv().roots().real)
        xn = 0.5 * (np.concatenate([[-1.0], inner, [1.0]]) + 1.0)
    else:
        xn = T["xn"]
    npn = order + 1
    j, i = np.meshgrid(np.arange(npn), np.arange(npn), indexing="ij")
    xi = np.stack([xn[i.ravel()], xn[j.ravel()]], 1)
    hv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], float)
    shp = np.ones((npn * npn, 4))
    for a in range(4):
        for d in range(2):
            shp[:, a] *= np.where(hv[a, d] > 0, xi[:, d], 1 - xi[:, d])
    xy = np.einsum("na,ead->end", shp, gm["elem_xyz"]).reshape(-1, 2)
    dof = npn * npn
    if mixture is None:
        Ug = ac.dry_state(xy, nvel)
        neq = nvel + 2
    else:
        up = ac.argon6_primitives(xy, nvel)
        import oracle_api
        pm = tps_b200.PlasmaModels.from_dict(mixture)
        neq = up.shape[1]
        orc = oracle_api.Oracle(order, gm["elem_xyz"], gm["face_el1"], gm["face_el2"], gm["face_inf1"], gm["face_inf2"],
                                phys=oracle_api.mixture_params(pm, 1), kind="ref", basis_type=bt, int_rule=ir, neq=neq, nvel=nvel)
        Ug = np.ascontiguousarray(orc.pt("cons", up).T.reshape(-1))
    NEg = gm["elem_xyz"].shape[0]
    distg = None
    if ml is not None:  # nodal wall distance: distance to the x = lo side
        distg = np.ascontiguousarray(np.abs(xy[:, 0] - 0.5) + 0.01)
        dg = torch.from_numpy(distg).to(dev)
        gop.set_distance_field(dg)
    yg = gop.Mult(torch.from_numpy(Ug).to(dev)).cpu().numpy().reshape(neq, NEg, dof)
    mcs_g = gop.max_char_speed()
    # partitioned operator
    specs = ac.bcs(bck, nvel, (0.02 * (ac.MW_AR - ac.MW_E), 0.05 * ac.MW_AR, 0.03 * ac.MW_AR, 0.04 * ac.MW_AR, 0.02 * ac.MW_E)
                   if mixture is not None else ())
    if mixture is not None:
        phys = tps_b200.Physics.plasma_mixture(tps_b200.PlasmaModels.from_dict(mixture), 1)
    else:
        phys = tps_b200.Physics.dry_air(1, 3e4, 0.2)
    if ml is not None:
        phys.with_mixing_length(*ml)
    op = tps_b200.RhsOperator(part, order=order, physics=phys, basis_type=bt, int_rule_type=ir, nvel=nvel, device=local_rank,
                              face_attr=part["face_attr"], use_bc_in_grad=ubg, bcs=[tps_b200.BcDesc.make(*b) for b in specs],
                              halo=tps_b200.make_halo_desc(part, comm), num_nbr_elems=part["num_nbr_elems"])
    ne = part["num_elems"]
    gid = part["elem_gid"][:ne]
    Ul = np.ascontiguousarray(Ug.reshape(neq, NEg, dof)[:, gid, :]).reshape(-1)
    if distg is not None:
        dl = torch.from_numpy(np.ascontiguousarray(distg.reshape(NEg, dof)[gid].reshape(-1))).to(dev)
        op.set_distance_field(dl)
    x = torch.from_numpy(Ul).to(dev)
    for it in range(3):
        y = op.Mult(x).cpu().numpy().reshape(neq, ne, dof)
    errs = [rel_l2(y[k], yg[k][gid]) for k in range(neq)]
    mcs = op.max_char_speed()
    xs = torch.from_numpy(Ul.copy()).to(dev)
    op.ode_step(xs, 1e-7, scheme=4, nsteps=3)
    xg = torch.from_numpy(Ug.copy()).to(dev)
    gop.ode_step(xg, 1e-7, scheme=4, nsteps=3)
    err_rk = rel_l2(xs.cpu().numpy().reshape(neq, ne, dof), xg.cpu().numpy().reshape(neq, NEg, dof)[:, gid, :])
    print(f"[{case} {method}] rank {rank}/{world}: path={op.path()} ne={ne} halo={part['num_nbr_elems']} "
          f"max_eq rel_l2(N-rank vs 1-rank)={max(errs):.3e} rk4={err_rk:.3e} mcs {mcs:.12e} vs {mcs_g:.12e}", flush=True)
    ok = max(errs) < 1e-12 and err_rk < 1e-12 and abs(mcs / mcs_g - 1) < 1e-14
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
