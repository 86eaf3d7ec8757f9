"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.
Tolerances are BASELINE.json's: per-evaluation du/dt relative L2 <= 1e-10, solution after 100 RK steps
<= 1e-8; index maps bit-exact."""
import numpy as np
import pytest

import meshref
import oracle_api
import tps_b200
from common import node_coords_from_mesh, rel_l2, tgv_state, warp_mesh

pytestmark = pytest.mark.gpu
PI = np.pi


def _setup(n, order=3, eq=1, visc_mult=1.0, bulk=0.0, lo=(-PI,) * 3, hi=(PI,) * 3, order_mode=0, kind="port"):
    import torch
    n3 = (n, n, n) if isinstance(n, int) else n
    m = tps_b200.cartesian_hex_mesh(*n3, lo=lo, hi=hi, order_mode=order_mode)
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(eq, visc_mult, bulk))
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(eq, visc_mult, bulk), kind=kind)
    U = tgv_state(orc.node_coords())
    return torch, m, op, orc, U


def test_index_maps_bit_exact(lib_built, oracle_built):
    _, m, op, orc, _ = _setup(4)
    ref = meshref.element_to_faces(op.NE, m["face_el1"], m["face_el2"])
    assert np.array_equal(op.element_to_faces(), ref)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_primitives_and_gradients(lib_built, oracle_built, order):
    torch, m, op, orc, U = _setup((4, 3, 5), order=order)
    x = torch.from_numpy(U).cuda()
    op.updatePrimitives(x)
    op.updateGradients(x, True)
    up, g = op.fields()
    assert rel_l2(up.cpu().numpy(), orc.primitives(U)) < 1e-14
    assert rel_l2(g.cpu().numpy(), orc.gradients(U)) < 1e-11


@pytest.mark.parametrize("order,eq,vm", [(3, 1, 1.0), (3, 1, 5e4), (3, 0, 1.0), (2, 1, 5e4), (1, 1, 5e4)])
def test_rhs_mult_parity(lib_built, oracle_built, order, eq, vm):
    torch, m, op, orc, U = _setup((6, 6, 6) if order == 3 else (4, 5, 3), order=order, eq=eq, visc_mult=vm, bulk=0.3)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    assert abs(op.max_char_speed() / orc.max_char_speed - 1) < 1e-13


def test_rhs_mult_parity_against_reference_object_code(lib_built, oracle_built):
    """Same check with the oracle's per-point physics served by the reference's own compiled classes."""
    import os
    if not os.path.exists(os.path.join(oracle_api.ORACLE_DIR, "_ref", "liboracle_ref.so")):
        pytest.skip("oracle/_ref not built")
    torch, m, op, orc, U = _setup(5, visc_mult=2e4, kind="ref")
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    assert rel_l2(y, orc.mult(U)) < 1e-10


def test_stretched_box_and_blocked_element_order(lib_built, oracle_built):
    """Anisotropic (still affine) elements, and the locality-preserving element numbering used by bench.py."""
    torch, m, op, orc, U = _setup((9, 10, 11), visc_mult=1e4, lo=(0.0, -1.0, 2.0), hi=(3.0, 1.5, 2.7), order_mode=1)
    y = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    assert rel_l2(y, orc.mult(U)) < 1e-10


@pytest.mark.parametrize("order,eq", [(3, 1), (2, 1), (3, 0), (1, 1)])
def test_trilinear_mesh_parity(lib_built, oracle_built, order, eq):
    """Genuinely trilinear (non-parallelepiped) hexahedra: per-node Jacobians, face normals varying over the
    face -- the general kernels (the cyl3d-type meshes of BASELINE config C2 take this path)."""
    import torch
    m = warp_mesh(tps_b200.cartesian_hex_mesh(5, 4, 6, lo=(-PI,) * 3, hi=(PI,) * 3), amp=0.12)
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(eq, 3e4, 0.2))
    orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                            phys=oracle_api.dry_air_params(eq, 3e4, 0.2))
    U = tgv_state(orc.node_coords())
    x = torch.from_numpy(U).cuda()
    y = op.Mult(x).cpu().numpy()
    yo = orc.mult(U)
    N = orc.N
    g = op.fields()[1].cpu().numpy()
    assert rel_l2(g, orc.gradients(U)) < 1e-11
    for k in range(5):
        assert rel_l2(y[k * N:(k + 1) * N], yo[k * N:(k + 1) * N]) < 1e-10, k
    assert abs(op.max_char_speed() / orc.max_char_speed - 1) < 1e-13


def test_general_path_matches_fast_path_on_affine_mesh(lib_built, oracle_built, monkeypatch):
    """The two independent kernel sets (rhs_fast.cuh / general) agree to round-off where both apply."""
    torch, m, op, orc, U = _setup((5, 6, 4), visc_mult=2e4, bulk=0.1)
    y_fast = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    monkeypatch.setenv("TPSB_PATH", "general")
    op2 = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 2e4, 0.1))
    y_gen = op2.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    assert rel_l2(y_fast, y_gen) < 1e-12
    assert rel_l2(y_gen, orc.mult(U)) < 1e-10


def test_host_buffer_entry_point_matches_device_entry_point(lib_built, oracle_built):
    torch, m, op, orc, U = _setup(4, visc_mult=1e4)
    y_dev = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    hx = torch.from_numpy(U).pin_memory()
    hy = torch.empty_like(hx).pin_memory()
    op.mult_host(hx, hy)
    assert np.array_equal(hy.numpy(), y_dev)
    y2 = np.zeros_like(U)
    op.mult_host(U, y2)  # pageable numpy buffers
    assert np.array_equal(y2, y_dev)


@pytest.mark.parametrize("n3,chunks,order", [((6, 6, 6), 4, 3), ((5, 4, 7), 7, 3), ((8, 8, 8), 32, 2), ((4, 4, 4), 3, 1)])
def test_chunked_host_pipeline_is_bit_identical(lib_built, monkeypatch, n3, chunks, order):
    """tpsb_rhs_mult_host overlaps copy-in / kernels / copy-out per element chunk (periodic single-rank fast path);
    the chunked schedule must reproduce the one-shot evaluation bit for bit, max characteristic speed included."""
    import torch
    monkeypatch.setenv("TPSB_HOST_CHUNKS", str(chunks))
    m = tps_b200.cartesian_hex_mesh(*n3, lo=(-PI,) * 3, hi=(PI,) * 3)
    op = tps_b200.RhsOperator(m, order=order, physics=tps_b200.Physics.dry_air(1, 3e3, 0.2))
    U = tgv_state(node_coords_from_mesh(m["elem_xyz"], order))
    y_dev = op.Mult(torch.from_numpy(U).cuda()).cpu().numpy()
    mcs = op.max_char_speed()
    hx = torch.from_numpy(U).pin_memory()
    for _ in range(2):  # second call reuses the streams / events
        hy = torch.zeros_like(hx).pin_memory()
        op.mult_host(hx, hy)
        assert np.array_equal(hy.numpy(), y_dev)
        assert op.max_char_speed() == mcs
    y2 = np.zeros_like(U)
    op.mult_host(U, y2)  # pageable buffers
    assert np.array_equal(y2, y_dev)


def test_hundred_rk4_steps(lib_built, oracle_built):
    """Solution after 100 RK steps agrees to 1e-8 (BASELINE.json north star)."""
    torch, m, op, orc, U = _setup(4, visc_mult=1e3)
    orc.mult(U)
    h = 2 * PI / 4
    dt = 0.3 * (h / 3) / orc.max_char_speed / 3  # CFL-like, cf. src/M2ulPhyS.cpp:2014
    x = torch.from_numpy(U.copy()).cuda()
    op.ode_step(x, dt, scheme=4, nsteps=100)
    ref = orc.rk4(U, dt, 100)
    got = x.cpu().numpy()
    N = orc.N
    assert np.isfinite(got).all()
    assert rel_l2(ref, U) > 1e-5  # the state really moved
    for k in range(5):
        assert rel_l2(got[k * N:(k + 1) * N], ref[k * N:(k + 1) * N]) < 1e-8, k


@pytest.mark.parametrize("scheme", [1, 2, 3])
def test_other_ode_schemes_are_consistent(lib_built, oracle_built, scheme):
    """ForwardEuler / RK2 / RK3SSP against a numpy restatement driven by the CUDA Mult itself."""
    torch, m, op, orc, U = _setup(3, visc_mult=1e3)
    dt = 1e-5
    x = torch.from_numpy(U.copy()).cuda()
    op.ode_step(x, dt, scheme=scheme, nsteps=2)
    f = lambda v: op.Mult(torch.from_numpy(np.ascontiguousarray(v)).cuda()).cpu().numpy()
    u = U.copy()
    for _ in range(2):
        if scheme == 1:
            u = u + dt * f(u)
        elif scheme == 2:
            k1 = f(u)
            k2 = f(u + dt * k1)
            u = u + 0.5 * dt * (k1 + k2)
        else:
            y1 = u + dt * f(u)
            y2 = 0.75 * u + 0.25 * (y1 + dt * f(y1))
            u = u / 3.0 + 2.0 / 3.0 * (y2 + dt * f(y2))
    assert rel_l2(x.cpu().numpy(), u) < 1e-13


def test_full_size_properties(lib_built):
    """Size-independent properties at a BASELINE-scale mesh (64^3 hexes, 16.8 M nodes): a uniform state has
    zero residual, and on a periodic box the DG scheme is conservative: sum_nodes w|J| dU/dt = 0."""
    import torch
    n = 64
    m = tps_b200.cartesian_hex_mesh(n, n, n, lo=(-PI,) * 3, hi=(PI,) * 3, order_mode=1)
    op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 1e3))
    N = op.N
    U = torch.empty(5 * N, dtype=torch.float64, device="cuda")
    for k, v in enumerate((1.2, 12.0, -3.0, 5.0, 253000.0)):
        U[k * N:(k + 1) * N] = v
    y = op.Mult(U)
    assert float(y[:N].abs().max()) < 1e-7 and float(y[4 * N:].abs().max()) / 253000.0 < 1e-7
    xyz = node_coords_from_mesh(m["elem_xyz"], 3)
    Ut = torch.from_numpy(tgv_state(xyz)).cuda()
    y = op.Mult(Ut)
    from tps_b200 import capi
    T = capi.ref_tables(3)
    w1 = torch.from_numpy(T["wn"]).cuda()
    w = (w1[:, None, None] * w1[None, :, None] * w1[None, None, :]).reshape(-1) * (2 * PI / n) ** 3
    for k in range(5):
        yk = y[k * N:(k + 1) * N].reshape(-1, 64)
        total = float((yk * w[None, :]).sum())
        scale = float((yk.abs() * w[None, :]).sum())
        assert abs(total) < 1e-9 * scale, (k, total, scale)


@pytest.mark.parametrize("scheme", [1, 2, 3, 4])
def test_graph_replayed_ode_steps_equal_eager_steps(lib_built, monkeypatch, scheme):
    """tpsb_ode_step replays one captured Runge-Kutta step as a CUDA graph from the second step on; the replay must
    reproduce eager stepping bit for bit, also when dt / the solution vector change between calls (re-capture) and on
    a side stream."""
    import torch
    m = tps_b200.cartesian_hex_mesh(4, 3, 3, lo=(-PI,) * 3, hi=(PI,) * 3)
    U = tgv_state(node_coords_from_mesh(m["elem_xyz"], 3))

    def run(graph, stream=None):
        monkeypatch.setenv("TPSB_ODE_GRAPH", "1" if graph else "0")
        op = tps_b200.RhsOperator(m, order=3, physics=tps_b200.Physics.dry_air(1, 2e3, 0.1), stream=stream)
        x = torch.from_numpy(U.copy()).cuda()
        n0 = op.launch_count()
        op.ode_step(x, 2e-6, scheme=scheme, nsteps=7)
        n1 = op.launch_count()
        op.ode_step(x, 1e-6, scheme=scheme, nsteps=5)   # new dt: the graph is re-captured
        x2 = x.clone()
        op.ode_step(x2, 1e-6, scheme=scheme, nsteps=4)  # new solution vector
        torch.cuda.synchronize()
        return x.cpu().numpy(), x2.cpu().numpy(), n1 - n0

    e1, e2, ne = run(False)
    g1, g2, ng = run(True)
    assert np.array_equal(e1, g1) and np.array_equal(e2, g2)
    assert ne == ng  # the launch counter counts replayed kernels too
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        s1, s2, _ = run(True, stream=s.cuda_stream)
    assert np.array_equal(e1, s1) and np.array_equal(e2, s2)
