"""ctypes access to the CPU oracle (oracle/liboracle.so, oracle/_ref/liboracle_ref.so).
TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class OrcPhysParams(C.Structure):
    _fields_ = [("eq_system", C.c_int), ("fluid", C.c_int), ("gamma", C.c_double), ("R", C.c_double),
                ("visc_mult", C.c_double), ("bulk_visc_mult", C.c_double), ("C1", C.c_double),
                ("S0", C.c_double), ("Pr", C.c_double), ("plasma", C.c_void_p), ("use_roe", C.c_int),
                ("sgs_model", C.c_int), ("sgs_const", C.c_double), ("sgs_floor", C.c_double), ("sponge_enabled", C.c_int),
                ("sponge_normal", C.c_double * 3), ("sponge_point", C.c_double * 3), ("sponge_ratio", C.c_double),
                ("sponge_width", C.c_double), ("use_mixing_length", C.c_int), ("max_mixing_length", C.c_double),
                ("mixing_length_Prt", C.c_double), ("mixing_length_bulk_mult", C.c_double), ("lte", C.c_void_p)]


class OrcBc(C.Structure):
    """kind 0 inlet / 1 outlet / 2 wall; type = the reference's InletType / OutletType / WallType value."""
    _fields_ = [("attr", C.c_int), ("kind", C.c_int), ("type", C.c_int), ("data", C.c_double * 12)]


def make_bc(attr, kind, type_, data=()):
    b = OrcBc(attr, kind, type_)
    for i, v in enumerate(data):
        b.data[i] = float(v)
    return b


def lte_params(tables, eq_system=1):
    """LTE_FLUID over the 1-D tables of a tps_b200.LteTables (same layout as OrcLte); reference back end (kind="ref") only."""
    p = OrcPhysParams(eq_system, 2, 1.4, 287.058, 1.0, 0.0, 1.458e-6, 110.4, 0.71, None, 0)
    p.lte = C.addressof(tables)
    p._tables = tables
    return p


def dry_air_params(eq_system=1, visc_mult=1.0, bulk_visc_mult=0.0, use_roe=False, sgs=None, sponge=None):
    """Defaults of the reference: gamma/R src/equation_of_state.cpp:175-179; Sutherland SURVEY.md 8(d).
    sgs = (model, constant, floor); sponge = (normal, point, ratio, width) -- reference back end only."""
    p = OrcPhysParams(eq_system, 0, 1.4, 287.058, visc_mult, bulk_visc_mult, 1.458e-6, 110.4, 0.71, None, int(use_roe))
    set_visc_mods(p, sgs, sponge)
    return p


def set_visc_mods(p, sgs=None, sponge=None):
    """Fill the SGS / viscous-sponge fields shared by OrcPhysParams and tps_b200.Physics."""
    if sgs:
        p.sgs_model, p.sgs_const, p.sgs_floor = int(sgs[0]), float(sgs[1]), float(sgs[2])
    if sponge:
        p.sponge_enabled = 1
        for d in range(3):
            p.sponge_normal[d], p.sponge_point[d] = float(sponge[0][d]), float(sponge[1][d])
        p.sponge_ratio, p.sponge_width = float(sponge[2]), float(sponge[3])


def mixture_params(models, eq_system=1):
    """fluid = user_defined: `models` is a tps_b200.PlasmaModels (same layout as the oracle's OrcPlasma); only the
    reference-object-code flavour of the oracle (kind='ref') serves mixtures."""
    p = OrcPhysParams(eq_system, 1, 1.4, 287.058, 1.0, 0.0, 1.458e-6, 110.4, 0.71, C.addressof(models), 0)
    p._models = models
    return p


def build(ref=False):
    """Compile the oracle (and, when /root/reference is present, the reference-physics flavour)."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "all"])
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def load(kind="port"):
    path = os.path.join(ORACLE_DIR, "liboracle.so" if kind == "port" else "_ref/liboracle_ref.so")
    if not os.path.exists(path):
        if kind == "port":
            build()
        else:
            raise FileNotFoundError(path)
    lib = C.CDLL(path)
    lib.orc_physics_kind.restype = C.c_char_p
    lib.orc_create.restype = C.c_void_p
    lib.orc_create.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _ip, _ip, _ip, _ip, C.POINTER(OrcPhysParams), C.c_int]
    lib.orc_create_ex.restype = C.c_void_p
    lib.orc_create_ex.argtypes = [C.c_int] * 7 + [_dp, C.c_int, _ip, _ip, _ip, _ip, C.POINTER(OrcPhysParams), C.c_int]
    lib.orc_destroy.argtypes = [C.c_void_p]
    lib.orc_set_bcs.argtypes = [C.c_void_p, _ip, C.c_int, C.POINTER(OrcBc), C.c_int]
    lib.orc_bc_flux.argtypes = [C.c_void_p, C.POINTER(OrcBc), C.c_int, _dp, _dp, _dp, _dp]
    lib.orc_set_bc_time_step.argtypes = [C.c_void_p, C.c_double]
    lib.orc_get_bc_state.argtypes = [C.c_void_p, C.c_int, _dp, _dp, C.c_int]
    lib.orc_get_bc_state.restype = C.c_int
    lib.orc_set_solution_view.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_add_forcing.argtypes = [C.c_void_p, C.c_void_p]
    lib.orc_clear_forcings.argtypes = [C.c_void_p]
    lib.orc_set_rates.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    for nm in ("orc_pt_prim", "orc_pt_cons", "orc_pt_max_char_speed", "orc_pt_conv_flux"):
        getattr(lib, nm).argtypes = [C.c_void_p, C.c_int, _dp, _dp]
    lib.orc_pt_visc_flux.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
    lib.orc_pt_source.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, _dp]
    lib.orc_pt_mix_diffusivity.argtypes = [C.c_void_p, _dp, _dp]
    lib.orc_ndofs.restype = C.c_long
    lib.orc_ndofs.argtypes = [C.c_void_p]
    lib.orc_update_primitives.argtypes = [C.c_void_p, _dp, _dp]
    lib.orc_compute_gradients.argtypes = [C.c_void_p, _dp, _dp]
    lib.orc_rhs_mult.argtypes = [C.c_void_p, _dp, _dp, C.c_void_p, C.POINTER(C.c_double)]
    lib.orc_rk4_steps.argtypes = [C.c_void_p, _dp, C.c_double, C.c_int]
    lib.orc_node_coords.argtypes = [C.c_void_p, _dp]
    lib.orc_face_nq.argtypes = [C.c_void_p]
    lib.orc_face_geometry.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
    lib.orc_elem_size.argtypes = [C.c_void_p, _dp]
    lib.orc_set_distance.argtypes = [C.c_void_p, _dp]
    lib.orc_dense_ops.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orc_gl_rule.argtypes = [C.c_int, _dp, _dp]
    lib.orc_phys_init.argtypes = [C.POINTER(OrcPhysParams)]
    lib.orc_phys_prim.argtypes = [C.c_int, _dp, _dp]
    lib.orc_phys_max_char_speed.argtypes = [C.c_int, _dp, _dp]
    lib.orc_phys_conv_flux.argtypes = [C.c_int, _dp, _dp]
    lib.orc_phys_visc_flux.argtypes = [C.c_int, _dp, _dp, _dp]
    lib.orc_phys_riemann.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
    return lib


class Oracle:
    """Thin object over orc_*: one DG operator on one mesh."""

    def __init__(self, order, elem_xyz, el1, el2, inf1, inf2, phys=None, nthreads=None, kind="port", basis_type=0,
                 int_rule=0, neq=None, nvel=None):
        self.lib = load(kind)
        self.phys = phys or dry_air_params()
        self.NE = elem_xyz.shape[0]
        self.order = order
        self.dim = elem_xyz.shape[2]  # [NE][2^dim][dim]
        self.nvel = nvel or self.dim
        self.neq = neq or self.nvel + 2
        nthreads = nthreads or os.cpu_count()
        self.h = self.lib.orc_create_ex(self.dim, order, basis_type, int_rule, self.neq, self.nvel, self.NE,
                                        np.ascontiguousarray(elem_xyz, dtype=np.float64), len(el1),
                                        np.ascontiguousarray(el1, np.int32), np.ascontiguousarray(el2, np.int32),
                                        np.ascontiguousarray(inf1, np.int32), np.ascontiguousarray(inf2, np.int32),
                                        C.byref(self.phys), nthreads)
        assert self.h, "orc_create failed"
        self.N = self.lib.orc_ndofs(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def set_distance(self, dist):
        """Nodal wall distance read by the mixing-length model (kind='ref')."""
        self._dist = np.ascontiguousarray(dist, np.float64)
        self.lib.orc_set_distance(self.h, self._dist)

    def set_bcs(self, face_attr, bcs, use_bc_in_grad=False):
        """bcs: list of OrcBc (make_bc); face_attr: boundary attribute per face (0 on interior faces)."""
        arr = (OrcBc * len(bcs))(*bcs)
        self._bcs = arr
        self.lib.orc_set_bcs(self.h, np.ascontiguousarray(face_attr, np.int32), len(bcs), arr, int(use_bc_in_grad))

    def set_bc_time_step(self, dt):
        self.lib.orc_set_bc_time_step(self.h, float(dt))

    def bc_state(self, attr):
        mean = np.zeros(self.neq)
        n = self.lib.orc_get_bc_state(self.h, int(attr), mean, np.zeros(1), 0)
        assert n >= 0, "no non-reflecting condition on this attribute"
        bu = np.zeros((max(n, 1), self.neq))
        self.lib.orc_get_bc_state(self.h, int(attr), mean, bu, n)
        return mean, bu[:n]

    def bc_flux(self, bc, normal, state, grad, use_bc_in_grad=False):
        out = np.zeros(self.neq)
        self.lib.orc_bc_flux(self.h, C.byref(bc), int(use_bc_in_grad), np.ascontiguousarray(normal, dtype=np.float64),
                             np.ascontiguousarray(state, dtype=np.float64), np.ascontiguousarray(grad, dtype=np.float64), out)
        return out

    def add_forcing(self, desc):
        """desc: a tps_b200.capi.ForcingDesc (same layout as the oracle's OrcForcing); a Joule-heating field must be a
        host array kept alive by the caller."""
        self._forcings = getattr(self, "_forcings", []) + [desc]
        self.lib.orc_add_forcing(self.h, C.addressof(desc))

    def set_solution_view(self, U):
        self._sol = None if U is None else np.ascontiguousarray(U, dtype=np.float64)
        self.lib.orc_set_solution_view(self.h, None if U is None else self._sol.ctypes.data)

    def set_rates(self, rates):
        """rates[component][N]: rate coefficients of the GRIDFUNCTION_RXN reactions (Chemistry::setGridFunctionRates)."""
        self._rates = np.ascontiguousarray(rates, dtype=np.float64)
        self.lib.orc_set_rates(self.h, self._rates.ctypes.data, self._rates.shape[-1])

    # point-wise probes of this operator's physics object: arrays are [n][neq] / [n][dim][neq] point-major
    def pt(self, what, *arrs):
        a0 = np.ascontiguousarray(arrs[0], dtype=np.float64)
        n = a0.shape[0]
        shape = {"prim": (n, self.neq), "cons": (n, self.neq), "max_char_speed": (n,), "conv_flux": (n, self.dim * self.neq),
                 "visc_flux": (n, self.dim * self.neq), "source": (n, self.neq)}[what]
        out = np.zeros(shape)
        args = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
        getattr(self.lib, "orc_pt_" + what)(self.h, n, *args, out)
        return out

    def mixture_average_diffusivity(self, state, num_species):
        """MolecularTransport::computeMixtureAverageDiffusivity of the reference's transport object (kind='ref')."""
        D = np.zeros(max(num_species, 8))
        rc = self.lib.orc_pt_mix_diffusivity(self.h, np.ascontiguousarray(state, dtype=np.float64), D)
        assert rc == 0, "this physics back end has no collision-integral transport"
        return D[:num_species]

    def node_coords(self):
        xyz = np.zeros((self.N, self.dim))
        self.lib.orc_node_coords(self.h, xyz)
        return xyz

    def primitives(self, x):
        up = np.zeros_like(x)
        self.lib.orc_update_primitives(self.h, x, up)
        return up

    def gradients(self, x):
        g = np.zeros(self.N * self.neq * self.dim)
        self.lib.orc_compute_gradients(self.h, x, g)
        return g

    def mult(self, x, want_grad=False):
        y = np.zeros_like(x)
        g = np.zeros(self.N * self.neq * self.dim) if want_grad else None
        mcs = C.c_double(0.0)
        self.lib.orc_rhs_mult(self.h, x, y, g.ctypes.data if want_grad else None, C.byref(mcs))
        self.max_char_speed = mcs.value
        return (y, g) if want_grad else y

    def rk4(self, U, dt, nsteps):
        U = np.ascontiguousarray(U.copy())
        self.lib.orc_rk4_steps(self.h, U, dt, nsteps)
        return U

    def face_geometry(self, f):
        nq = self.lib.orc_face_nq(self.h)
        nor, xyz, w = np.zeros((nq, self.dim)), np.zeros((nq, self.dim)), np.zeros(nq)
        self.lib.orc_face_geometry(self.h, f, nor, xyz, w)
        return nor, xyz, w
