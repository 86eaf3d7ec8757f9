"""ctypes binding of include/tpsb200.h.  Device memory and streams come from torch (plumbing only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class TpsbError(RuntimeError):
    pass


def library_path():
    return os.path.join(_HERE, "lib", "libtpsb200.so")


def build_library():
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "csrc")])
    return library_path()


class MeshMaps(C.Structure):
    _fields_ = [("dim", C.c_int), ("num_elems", C.c_int), ("num_nbr_elems", C.c_int),
                ("elem_vertices", C.POINTER(C.c_double)), ("num_faces", C.c_int),
                ("face_el1", C.POINTER(C.c_int)), ("face_el2", C.POINTER(C.c_int)),
                ("face_inf1", C.POINTER(C.c_int)), ("face_inf2", C.POINTER(C.c_int)),
                ("face_attr", C.POINTER(C.c_int))]


class SpaceDesc(C.Structure):
    _fields_ = [("order", C.c_int), ("basis_type", C.c_int), ("int_rule_type", C.c_int),
                ("num_equation", C.c_int), ("nvel", C.c_int)]


MAX_SPECIES, MAX_REACTIONS = 8, 34


class PlasmaModels(C.Structure):
    """tpsb_plasma_models: PerfectMixtureInput + constantTransportData + ChemistryInput flattened, species in the
    reference's mixture order (electron second to last, background last)."""
    _fields_ = [("num_species", C.c_int), ("ambipolar", C.c_int), ("two_temperature", C.c_int),
                ("mw", C.c_double * MAX_SPECIES), ("charge", C.c_double * MAX_SPECIES),
                ("formation_energy", C.c_double * MAX_SPECIES), ("molar_cv", C.c_double * MAX_SPECIES),
                ("transport_model", C.c_int), ("viscosity", C.c_double), ("bulk_viscosity", C.c_double),
                ("thermal_conductivity", C.c_double), ("electron_thermal_conductivity", C.c_double),
                ("diffusivity", C.c_double * MAX_SPECIES), ("mt_freq", C.c_double * MAX_SPECIES),
                ("num_reactions", C.c_int), ("min_temperature", C.c_double),
                ("model", C.c_int * MAX_REACTIONS), ("detailed_balance", C.c_int * MAX_REACTIONS),
                ("rate_params", (C.c_double * 3) * MAX_REACTIONS), ("reaction_energy", C.c_double * MAX_REACTIONS),
                ("equilibrium_params", (C.c_double * 3) * MAX_REACTIONS),
                ("reactant_stoich", (C.c_int * MAX_SPECIES) * MAX_REACTIONS),
                ("product_stoich", (C.c_int * MAX_SPECIES) * MAX_REACTIONS),
                ("third_order_k_electron", C.c_int), ("multiply", C.c_int), ("flux_trns_multiplier", C.c_double * 4),
                ("mf_freq_multiplier", C.c_double), ("diff_mult", C.c_double), ("mobil_mult", C.c_double),
                ("table_n", C.c_int * MAX_REACTIONS), ("table_xlog", C.c_int * MAX_REACTIONS),
                ("table_flog", C.c_int * MAX_REACTIONS), ("table_x", C.POINTER(C.c_double) * MAX_REACTIONS),
                ("table_f", C.POINTER(C.c_double) * MAX_REACTIONS), ("rate_component", C.c_int * MAX_REACTIONS),
                ("collision_index", C.c_int * (MAX_SPECIES * MAX_SPECIES)), ("ion_index", C.c_int), ("neutral_index", C.c_int),
                ("nec_table_n", C.c_int), ("nec_table_xlog", C.c_int), ("nec_table_flog", C.c_int),
                ("nec_table_x", C.POINTER(C.c_double)), ("nec_table_f", C.POINTER(C.c_double))]

    @classmethod
    def from_dict(cls, d):
        """d: species = list of dicts (mw, charge, formation_energy, molar_cv, diffusivity, mt_freq) in mixture
        order, transport scalars, reactions = list of dicts (model, A, b, E, energy, detailed, eqA, eqB, eqE,
        reactants, products)."""
        pm = cls()
        sp = d["species"]
        pm.num_species, pm.ambipolar, pm.two_temperature = len(sp), int(d["ambipolar"]), int(d["two_temperature"])
        for i, s_ in enumerate(sp):
            pm.mw[i], pm.charge[i], pm.formation_energy[i] = s_["mw"], s_["charge"], s_["formation_energy"]
            pm.molar_cv[i], pm.diffusivity[i], pm.mt_freq[i] = s_["molar_cv"], s_.get("diffusivity", 0.0), s_.get("mt_freq", 0.0)
        pm.transport_model = {"constant": 2, "argon_minimal": 0, "argon_mixture": 1}[d.get("transport_model", "constant")]
        if pm.transport_model == 1:
            # M2ulPhyS::identifyCollisionType (src/M2ulPhyS.cpp:3925-3970); species kinds 'ion' (Ar.+1), 'electron',
            # 'neutral' (Ar and its excited states)
            ns = len(sp)
            kinds = [("electron" if q["charge"] < 0 else "ion" if q["charge"] > 0 else "neutral") for q in sp]
            # nitrogen mixtures name each species' GasSpcs kind: 'N2', 'NI', 'NI1P', 'E' (the nitrogen branch of
            # identifyCollisionType, src/reactingFlow.cpp:3641-3676); GasColl values src/dataStructures.hpp:122-144
            nit = {("N2", "N2"): 12, ("N2", "NI"): 13, ("N2", "NI1P"): 10, ("E", "N2"): 11, ("NI", "NI"): 8, ("NI", "NI1P"): 6,
                   ("E", "NI"): 7}
            for i in range(ns):
                for j in range(i, ns):
                    zz = sp[i]["charge"] * sp[j]["charge"]
                    pair = {kinds[i], kinds[j]}
                    if zz == 0 and "kind" in sp[i]:
                        pm.collision_index[i + j * ns] = nit[tuple(sorted((sp[i]["kind"], sp[j]["kind"])))]
                        continue
                    pm.collision_index[i + j * ns] = (1 if zz > 0 else 0 if zz < 0 else 4 if pair == {"neutral"} else
                                                      2 if pair == {"neutral", "ion"} else 3)
            pm.ion_index = d.get("ion_index", kinds.index("ion"))
            pm.neutral_index = d.get("neutral_index", ns - 1)
        pm.third_order_k_electron = int(d.get("third_order_k_electron", False))
        mult = d.get("multipliers")
        pm.multiply = int(mult is not None)
        if mult is not None:
            for k, key in enumerate(("viscosity", "bulk_viscosity", "heavy_thermal_conductivity", "electron_thermal_conductivity")):
                pm.flux_trns_multiplier[k] = mult.get(key, 1.0)
            pm.mf_freq_multiplier, pm.diff_mult, pm.mobil_mult = mult.get("momentum_transfer_frequency", 1.0), mult.get("diffusivity", 1.0), mult.get("mobility", 1.0)
        pm.viscosity, pm.bulk_viscosity = d.get("viscosity", 0.0), d.get("bulk_viscosity", 0.0)
        pm.thermal_conductivity, pm.electron_thermal_conductivity = d.get("thermal_conductivity", 0.0), d.get("electron_thermal_conductivity", 0.0)
        if d.get("nec_table") is not None:  # net emission coefficient table (x, f, xlog, flog)
            tx, tf = (np.ascontiguousarray(t, dtype=np.float64) for t in d["nec_table"][:2])
            pm._keep = getattr(pm, "_keep", []) + [tx, tf]
            pm.nec_table_n, pm.nec_table_xlog, pm.nec_table_flog = len(tx), int(d["nec_table"][2]), int(d["nec_table"][3])
            pm.nec_table_x, pm.nec_table_f = _dp(tx), _dp(tf)
        rx = d.get("reactions", [])
        pm.num_reactions, pm.min_temperature = len(rx), d.get("min_temperature", 0.0)
        for r, q in enumerate(rx):
            pm.model[r], pm.detailed_balance[r], pm.reaction_energy[r] = q.get("model", 0), int(q["detailed"]), q["energy"]
            for k, key in enumerate(("A", "b", "E")):
                pm.rate_params[r][k] = q.get(key, 0.0)
            for k, key in enumerate(("eqA", "eqB", "eqE")):
                pm.equilibrium_params[r][k] = q.get(key, 0.0)
            for i in range(len(sp)):
                pm.reactant_stoich[r][i], pm.product_stoich[r][i] = q["reactants"][i], q["products"][i]
            if q.get("model", 0) == 2:  # tabulated: table = (x, f, xlog, flog)
                tx, tf = (np.ascontiguousarray(t, dtype=np.float64) for t in q["table"][:2])
                pm._keep = getattr(pm, "_keep", []) + [tx, tf]
                pm.table_n[r], pm.table_xlog[r], pm.table_flog[r] = len(tx), int(q["table"][2]), int(q["table"][3])
                pm.table_x[r], pm.table_f[r] = _dp(tx), _dp(tf)
            pm.rate_component[r] = q.get("component", 0)
        return pm


class LteTables(C.Structure):
    """tpsb_lte_tables (same layout as the oracle's OrcLte): 1-D look-up tables of the LTE working fluid."""
    _fields_ = [("num_thermo", C.c_int), ("T", C.c_void_p), ("energy", C.c_void_p), ("R", C.c_void_p), ("c", C.c_void_p),
                ("num_trans", C.c_int), ("T_trans", C.c_void_p), ("mu", C.c_void_p), ("kappa", C.c_void_p), ("sigma", C.c_void_p),
                ("nec_table_n", C.c_int), ("nec_table_xlog", C.c_int), ("nec_table_flog", C.c_int),
                ("nec_table_x", C.c_void_p), ("nec_table_f", C.c_void_p)]

    @classmethod
    def make(cls, T, energy, R, c, T_trans, mu, kappa, sigma, nec=None):
        """nec = (x, f, xlog, flog) of the net emission coefficient, or None.  The arrays are kept alive on the object."""
        t = cls()
        keep = [np.ascontiguousarray(a, dtype=np.float64) for a in (T, energy, R, c, T_trans, mu, kappa, sigma)]
        t.num_thermo, t.num_trans = len(keep[0]), len(keep[4])
        for name, a in zip(("T", "energy", "R", "c", "T_trans", "mu", "kappa", "sigma"), keep):
            setattr(t, name, a.ctypes.data)
        if nec is not None:
            nx, nf = np.ascontiguousarray(nec[0], dtype=np.float64), np.ascontiguousarray(nec[1], dtype=np.float64)
            keep += [nx, nf]
            t.nec_table_n, t.nec_table_xlog, t.nec_table_flog = len(nx), int(nec[2]), int(nec[3])
            t.nec_table_x, t.nec_table_f = nx.ctypes.data, nf.ctypes.data
        t._keep = keep
        return t


class Physics(C.Structure):
    """tpsb_physics; defaults are the reference's dry-air constants."""
    _fields_ = [("eq_system", C.c_int), ("fluid", C.c_int), ("specific_heat_ratio", C.c_double),
                ("gas_constant", C.c_double), ("visc_mult", C.c_double), ("bulk_visc_mult", C.c_double),
                ("sutherland_C1", C.c_double), ("sutherland_S0", C.c_double), ("sutherland_Pr", C.c_double),
                ("plasma", C.POINTER(PlasmaModels)), ("use_roe", C.c_int),
                ("sgs_model", C.c_int), ("sgs_const", C.c_double), ("sgs_floor", C.c_double), ("sponge_enabled", C.c_int),
                ("sponge_normal", C.c_double * 3), ("sponge_point", C.c_double * 3), ("sponge_ratio", C.c_double),
                ("sponge_width", C.c_double), ("use_mixing_length", C.c_int), ("max_mixing_length", C.c_double),
                ("mixing_length_Prt", C.c_double), ("mixing_length_bulk_mult", C.c_double), ("lte", C.POINTER(LteTables))]

    def with_mixing_length(self, max_mixing_length, pr_ratio=1.0, bulk_multiplier=0.0):
        """flow/useMixingLength = True with mixing-length/{max-mixing-length, Pr_ratio, bulk-multiplier}."""
        self.use_mixing_length, self.max_mixing_length = 1, float(max_mixing_length)
        self.mixing_length_Prt, self.mixing_length_bulk_mult = float(pr_ratio), float(bulk_multiplier)
        return self

    @classmethod
    def dry_air(cls, eq_system=1, visc_mult=1.0, bulk_visc_mult=0.0, use_roe=False, sgs=None, sponge=None):
        """sgs = (model 1 smagorinsky / 2 sigma, constant, floor); sponge = (normal, point, ratio, width)."""
        ph = cls(eq_system, 0, 1.4, 287.058, visc_mult, bulk_visc_mult, 1.458e-6, 110.4, 0.71, None, int(use_roe))
        if sgs:
            ph.sgs_model, ph.sgs_const, ph.sgs_floor = int(sgs[0]), float(sgs[1]), float(sgs[2])
        if sponge:
            ph.sponge_enabled = 1
            for d in range(3):
                ph.sponge_normal[d], ph.sponge_point[d] = float(sponge[0][d]), float(sponge[1][d])
            ph.sponge_ratio, ph.sponge_width = float(sponge[2]), float(sponge[3])
        return ph

    @classmethod
    def lte_fluid(cls, tables, eq_system=1):
        """fluid = LTE_FLUID with 1-D look-up tables (LteTables.make), kept alive on the returned object."""
        ph = cls(eq_system, 2, 1.4, 287.058, 1.0, 0.0, 1.458e-6, 110.4, 0.71, None, 0)
        ph.lte = C.pointer(tables)
        ph._tables = tables
        return ph

    @classmethod
    def plasma_mixture(cls, models, eq_system=1):
        """fluid = user_defined with the given PlasmaModels (kept alive on the returned object)."""
        ph = cls(eq_system, 1, 1.4, 287.058, 1.0, 0.0, 1.458e-6, 110.4, 0.71, C.pointer(models), 0)
        ph._models = models
        return ph


class BcDesc(C.Structure):
    """tpsb_bc_desc: kind 0 inlet / 1 outlet / 2 wall, type = the reference's InletType / OutletType / WallType."""
    _fields_ = [("attr", C.c_int), ("kind", C.c_int), ("type", C.c_int), ("data", C.c_double * 12)]

    @classmethod
    def make(cls, attr, kind, type_, data=()):
        b = cls(attr, kind, type_)
        for i, v in enumerate(data):
            b.data[i] = float(v)
        return b


class BcSet(C.Structure):
    _fields_ = [("num_bcs", C.c_int), ("bcs", C.POINTER(BcDesc)), ("use_bc_in_grad", C.c_int)]


class HaloDesc(C.Structure):
    _fields_ = [("num_nbr_ranks", C.c_int), ("nbr_rank", C.POINTER(C.c_int)), ("send_offset", C.POINTER(C.c_int)),
                ("send_elems", C.POINTER(C.c_int)), ("recv_offset", C.POINTER(C.c_int)), ("nccl_comm", C.c_void_p)]


class ForcingDesc(C.Structure):
    """tpsb_forcing_desc: kind 0 pressure gradient, 1 heat source (cylinder), 2 Joule heating (nodal field), 3 sponge zone."""
    _fields_ = [("kind", C.c_int), ("pressure_grad", C.c_double * 3), ("hs_point1", C.c_double * 3), ("hs_point2", C.c_double * 3),
                ("hs_radius", C.c_double), ("hs_value", C.c_double), ("joule_heating", C.c_void_p), ("sz_type", C.c_int),
                ("sz_mixed_out", C.c_int), ("sz_normal", C.c_double * 3), ("sz_point0", C.c_double * 3),
                ("sz_point_init", C.c_double * 3), ("sz_r1", C.c_double), ("sz_r2", C.c_double), ("sz_tol", C.c_double),
                ("sz_mult", C.c_double), ("sz_target", C.c_double * 5)]


class PartSizes(C.Structure):
    _fields_ = [("num_elems", C.c_int), ("num_nbr_elems", C.c_int), ("num_faces", C.c_int),
                ("num_nbr_ranks", C.c_int), ("num_send", C.c_int)]


# every symbol include/tpsb200.h declares (tests check the library exports all of them)
EXPORTS = ["tpsb_version", "tpsb_last_error", "tpsb_create", "tpsb_destroy", "tpsb_num_dofs", "tpsb_num_equation",
           "tpsb_rhs_mult", "tpsb_rhs_mult_host", "tpsb_update_primitives", "tpsb_update_gradients",
           "tpsb_get_fields", "tpsb_set_solution_view", "tpsb_set_reaction_rate_field", "tpsb_get_mean_time_derivatives", "tpsb_get_max_char_speed", "tpsb_ode_step", "tpsb_get_element_to_faces",
           "tpsb_launch_count", "tpsb_debug_buffer", "tpsb_debug_point_eval", "tpsb_set_distance_field", "tpsb_get_hmin", "tpsb_solve_step", "tpsb_check_state", "tpsb_debug_host_pipe_schedule", "tpsb_set_profiling", "tpsb_get_kernel_times", "tpsb_get_ref_tables", "tpsb_mk_cartesian_hex", "tpsb_mk_build_faces", "tpsb_mk_cartesian_quad", "tpsb_mk_build_faces2d", "tpsb_mk_partition", "tpsb_comm_get_unique_id",
           "tpsb_comm_init_rank", "tpsb_comm_destroy", "tpsb_get_path", "tpsb_mk_partition_metis", "tpsb_mk_partition_rcb",
           "tpsb_mk_partition_general", "tpsb_mk_partition_rcb_dim", "tpsb_mk_partition_general_dim", "tpsb_add_forcing", "tpsb_clear_forcings", "tpsb_averaging_add_sample",
           "tpsb_set_time_step", "tpsb_get_bc_state"]


def lib():
    """Load libtpsb200.so; raises loudly if the CUDA extension has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise TpsbError(f"{path} is missing: run tps_b200.build_library() / __graft_entry__.build(); "
                        "there is no CPU fallback for the RHS path")
    try:
        # torch bundles a newer libnccl.so.2 than the system one; whichever is loaded first wins the
        # soname, so let torch load its copy before libtpsb200.so pulls in NCCL.
        import torch  # noqa: F401
    except ImportError:
        pass
    L = C.CDLL(path)
    vp, ip, dp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)
    L.tpsb_version.restype = C.c_char_p
    L.tpsb_last_error.restype = C.c_char_p
    L.tpsb_last_error.argtypes = [vp]
    L.tpsb_create.argtypes = [C.POINTER(MeshMaps), C.POINTER(SpaceDesc), C.POINTER(Physics), C.POINTER(BcSet), C.POINTER(HaloDesc),
                              C.c_int, vp, C.POINTER(vp)]
    L.tpsb_destroy.argtypes = [vp]
    L.tpsb_destroy.restype = None
    L.tpsb_num_dofs.restype = C.c_int64
    L.tpsb_num_dofs.argtypes = [vp]
    L.tpsb_num_equation.argtypes = [vp]
    L.tpsb_add_forcing.argtypes = [vp, C.POINTER(ForcingDesc)]
    L.tpsb_clear_forcings.argtypes = [vp]
    L.tpsb_averaging_add_sample.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.tpsb_get_path.argtypes = [vp]
    L.tpsb_rhs_mult.argtypes = [vp, vp, vp]
    L.tpsb_rhs_mult_host.argtypes = [vp, vp, vp]
    L.tpsb_update_primitives.argtypes = [vp, vp]
    L.tpsb_update_gradients.argtypes = [vp, vp, C.c_int]
    L.tpsb_get_fields.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.tpsb_get_max_char_speed.argtypes = [vp, dp]
    L.tpsb_set_solution_view.argtypes = [vp, vp]
    L.tpsb_set_distance_field.argtypes = [vp, vp]
    L.tpsb_set_time_step.argtypes = [vp, C.c_double]
    L.tpsb_get_bc_state.argtypes = [vp, C.c_int, dp, dp, C.c_int, ip]
    L.tpsb_get_hmin.argtypes = [vp, dp]
    L.tpsb_check_state.argtypes = [vp, vp, ip]
    L.tpsb_solve_step.argtypes = [vp, vp, C.c_double, C.c_int, C.c_double, ip, dp]
    L.tpsb_debug_host_pipe_schedule.argtypes = [C.POINTER(MeshMaps), C.c_int, ip, ip, ip, ip, C.c_int, C.POINTER(C.c_int)]
    L.tpsb_set_reaction_rate_field.argtypes = [vp, vp, C.c_int]
    L.tpsb_get_mean_time_derivatives.argtypes = [vp, vp, dp]
    L.tpsb_ode_step.argtypes = [vp, vp, C.c_double, C.c_int, C.c_int]
    L.tpsb_get_element_to_faces.argtypes = [vp, ip]
    L.tpsb_debug_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64)]
    L.tpsb_debug_point_eval.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    L.tpsb_launch_count.restype = C.c_int64
    L.tpsb_launch_count.argtypes = [vp]
    L.tpsb_set_profiling.argtypes = [vp, C.c_int]
    L.tpsb_get_kernel_times.argtypes = [vp, dp, C.POINTER(C.c_int64)]
    L.tpsb_get_ref_tables.argtypes = [C.c_int, dp, C.c_int]
    L.tpsb_mk_cartesian_hex.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, ip, C.c_int, ip, dp]
    L.tpsb_mk_build_faces.argtypes = [C.c_int, ip, ip, ip, ip, ip]
    L.tpsb_mk_cartesian_quad.argtypes = [C.c_int, C.c_int, dp, dp, ip, ip, dp]
    L.tpsb_mk_build_faces2d.argtypes = [C.c_int, ip, ip, ip, ip, ip]
    L.tpsb_mk_partition.argtypes = [ip, dp, dp, ip, ip, C.c_int, C.c_int, C.POINTER(PartSizes), ip, dp,
                                    C.POINTER(C.c_int64), ip, ip, ip, ip, ip, ip, ip, ip]
    i64p = C.POINTER(C.c_int64)
    L.tpsb_mk_partition_metis.argtypes = [C.c_int, C.c_int, ip, ip, C.c_int, ip, i64p]
    L.tpsb_mk_partition_rcb.argtypes = [C.c_int, dp, C.c_int, ip]
    L.tpsb_mk_partition_general.argtypes = [C.c_int, ip, dp, C.c_int, ip, ip, ip, C.c_int, C.POINTER(PartSizes), ip, dp, i64p,
                                            ip, ip, ip, ip, ip, ip, ip, ip, ip]
    L.tpsb_mk_partition_general_dim.argtypes = [C.c_int] + L.tpsb_mk_partition_general.argtypes
    L.tpsb_mk_partition_rcb_dim.argtypes = [C.c_int, C.c_int, dp, C.c_int, ip]
    L.tpsb_comm_get_unique_id.argtypes = [C.c_char_p]
    L.tpsb_comm_init_rank.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.tpsb_comm_destroy.argtypes = [vp]
    _LIB = L
    return L


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def cartesian_hex_mesh(nx, ny, nz, lo=(-1.0, -1.0, -1.0), hi=(1.0, 1.0, 1.0), periodic=(1, 1, 1), order_mode=0):
    """meshkit: Cartesian hex box + MFEM-convention face tables (host only, no GPU needed)."""
    L = lib()
    NE = nx * ny * nz
    ev = np.zeros((NE, 8), dtype=np.int32)
    xyz = np.zeros((NE, 8, 3), dtype=np.float64)
    lo_a, hi_a = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    per = np.asarray(periodic, dtype=np.int32)
    rc = L.tpsb_mk_cartesian_hex(nx, ny, nz, _dp(lo_a), _dp(hi_a), _ip(per), order_mode, _ip(ev), _dp(xyz))
    if rc != 0:
        raise TpsbError(f"tpsb_mk_cartesian_hex failed with code {rc}")
    el1 = np.zeros(6 * NE, dtype=np.int32)
    el2, inf1, inf2 = np.zeros_like(el1), np.zeros_like(el1), np.zeros_like(el1)
    nf = L.tpsb_mk_build_faces(NE, _ip(ev), _ip(el1), _ip(el2), _ip(inf1), _ip(inf2))
    if nf < 0:
        raise TpsbError(f"tpsb_mk_build_faces failed with code {nf}")
    return dict(elem_verts=ev, elem_xyz=xyz, face_el1=el1[:nf].copy(), face_el2=el2[:nf].copy(),
                face_inf1=inf1[:nf].copy(), face_inf2=inf2[:nf].copy())


def cartesian_quad_mesh(nx, ny, lo=(-1.0, -1.0), hi=(1.0, 1.0), periodic=(1, 1)):
    """meshkit: Cartesian quadrilateral box + MFEM-convention edge tables (2-D test cases, utils/beam_mesh.cpp)."""
    L = lib()
    NE = nx * ny
    ev = np.zeros((NE, 4), dtype=np.int32)
    xyz = np.zeros((NE, 4, 2), dtype=np.float64)
    lo_a, hi_a = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    per = np.asarray(periodic, dtype=np.int32)
    rc = L.tpsb_mk_cartesian_quad(nx, ny, _dp(lo_a), _dp(hi_a), _ip(per), _ip(ev), _dp(xyz))
    if rc != 0:
        raise TpsbError(f"tpsb_mk_cartesian_quad failed with code {rc}")
    el1 = np.zeros(4 * NE, dtype=np.int32)
    el2, inf1, inf2 = np.zeros_like(el1), np.zeros_like(el1), np.zeros_like(el1)
    nf = L.tpsb_mk_build_faces2d(NE, _ip(ev), _ip(el1), _ip(el2), _ip(inf1), _ip(inf2))
    if nf < 0:
        raise TpsbError(f"tpsb_mk_build_faces2d failed with code {nf}")
    return dict(dim=2, elem_verts=ev, elem_xyz=xyz, face_el1=el1[:nf].copy(), face_el2=el2[:nf].copy(),
                face_inf1=inf1[:nf].copy(), face_inf2=inf2[:nf].copy())


QUAD_EDGE_VERT = np.array([[0, 1], [1, 2], [2, 3], [3, 0]])  # Geometry::Constants<SQUARE>::Edges


def quad_box_face_attr(m, lo, hi):
    """Boundary attribute of every face of a non-periodic Cartesian quad box: 1 x = lo, 2 x = hi, 3 y = lo, 4 y = hi
    (0 on interior faces)."""
    el1, el2, inf1 = m["face_el1"], m["face_el2"], m["face_inf1"]
    attr = np.zeros(len(el1), dtype=np.int32)
    tol = 1e-9 * max(hi[0] - lo[0], hi[1] - lo[1])
    for f in np.nonzero(el2 < 0)[0]:
        c = m["elem_xyz"][el1[f], QUAD_EDGE_VERT[inf1[f] // 64]].mean(axis=0)
        if abs(c[0] - lo[0]) < tol:
            attr[f] = 1
        elif abs(c[0] - hi[0]) < tol:
            attr[f] = 2
        elif abs(c[1] - lo[1]) < tol:
            attr[f] = 3
        elif abs(c[1] - hi[1]) < tol:
            attr[f] = 4
    return attr


HEX_FACE_VERT = np.array([[3, 2, 1, 0], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7], [4, 5, 6, 7]])


def cylinder_ogrid_mesh(nr, nth, nz, r_in=0.5, r_out=10.0, lz=2.0, stretch=1.08, order_mode=0):
    """meshkit: O-grid of trilinear hexahedra around a cylinder of diameter 2*r_in, extruded (periodic) in z --
    the hex restatement of the reference's cyl3d case (BASELINE config C2; its tet mesh is an LFS pointer).
    Built from the (r, theta, z) box: theta periodic, geometric radial stretching.  Returns the mesh dict plus
    'face_attr': 1 cylinder wall, 2 inlet (outer boundary, x < 0), 3 outlet (outer boundary, x >= 0)."""
    m = cartesian_hex_mesh(nr, nth, nz, lo=(0.0, 0.0, 0.0), hi=(1.0, 2 * np.pi, lz), periodic=(0, 1, 1),
                           order_mode=order_mode)
    s, th, z = m["elem_xyz"][..., 0], m["elem_xyz"][..., 1], m["elem_xyz"][..., 2]
    k = np.rint(s * nr)  # radial vertex index
    if abs(stretch - 1.0) < 1e-12:
        r = r_in + (r_out - r_in) * k / nr
    else:
        r = r_in + (r_out - r_in) * (stretch ** k - 1.0) / (stretch ** nr - 1.0)
    xyz = np.stack([r * np.cos(th), r * np.sin(th), z], axis=-1)
    el1, el2, inf1 = m["face_el1"], m["face_el2"], m["face_inf1"]
    attr = np.zeros(len(el1), dtype=np.int32)
    for f in np.nonzero(el2 < 0)[0]:
        fv = HEX_FACE_VERT[inf1[f] // 64]
        kk = k[el1[f], fv]
        if np.all(kk == 0):
            attr[f] = 1
        else:
            attr[f] = 2 if xyz[el1[f], fv, 0].mean() < 0 else 3
    m["elem_xyz"] = np.ascontiguousarray(xyz)
    m["face_attr"] = attr
    return m


def host_pipe_schedule(mesh, chunks, with_bdr=False):
    """Test hook: (elem_begin, face_begin, [(kind, chunk), ...]) of the chunked host-buffer pipeline, or None when the
    mesh cannot be chunked (tpsb_debug_host_pipe_schedule; host only)."""
    L = lib()
    xyz = np.ascontiguousarray(mesh["elem_xyz"], dtype=np.float64)
    arr = [np.ascontiguousarray(mesh[k], dtype=np.int32) for k in ("face_el1", "face_el2", "face_inf1", "face_inf2")]
    maps = MeshMaps(xyz.shape[2], xyz.shape[0], 0, _dp(xyz), len(arr[0]), _ip(arr[0]), _ip(arr[1]), _ip(arr[2]), _ip(arr[3]), None)
    eb, fb, bb = (np.zeros(chunks + 1, np.int32) for _ in range(3))
    ops, n = np.zeros(2 * 8 * max(chunks, 1), np.int32), C.c_int(0)
    rc = L.tpsb_debug_host_pipe_schedule(C.byref(maps), chunks, _ip(eb), _ip(fb), _ip(bb), _ip(ops), len(ops) // 2, C.byref(n))
    if rc != 0:
        return None
    sched = (eb, fb, [(int(ops[2 * k]), int(ops[2 * k + 1])) for k in range(n.value)])
    return sched + (bb,) if with_bdr else sched


def ref_tables(order):
    """Unpack tpsb_get_ref_tables into a dict (test hook)."""
    buf = np.zeros(4096)
    n = lib().tpsb_get_ref_tables(order, _dp(buf), len(buf))
    if n < 0:
        raise TpsbError(f"tpsb_get_ref_tables failed ({n})")
    np_, nq = int(buf[0]), int(buf[1])
    o = [2]

    def take(*shape):
        cnt = int(np.prod(shape))
        a = buf[o[0]:o[0] + cnt].reshape(shape).copy()
        o[0] += cnt
        return a
    T = dict(np=np_, nq=nq)
    T["xn"], T["wn"], T["D"], T["lb"] = take(np_), take(np_), take(np_, np_), take(2, np_)
    T["xq"], T["wq"], T["P"] = take(nq), take(nq), take(nq, np_)
    T["face_base"] = take(6, np_ * np_).astype(int)
    T["face_cstride"], T["face_side"] = take(6).astype(int), take(6).astype(int)
    T["perm"], T["iperm"] = take(8, np_ * np_).astype(int), take(8, np_ * np_).astype(int)
    return T


def cartesian_hex_partition(n, procs, rank, lo=(-1.0, -1.0, -1.0), hi=(1.0, 1.0, 1.0), periodic=(1, 1, 1),
                            order_mode=0):
    """meshkit: this rank's block of the box + face-neighbour (halo) tables (host only)."""
    L = lib()
    n_a, p_a = np.asarray(n, dtype=np.int32), np.asarray(procs, dtype=np.int32)
    lo_a, hi_a = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    per = np.asarray(periodic, dtype=np.int32)
    sz = PartSizes()
    args = (_ip(n_a), _dp(lo_a), _dp(hi_a), _ip(per), _ip(p_a), rank, order_mode, C.byref(sz))
    rc = L.tpsb_mk_partition(*args, None, None, None, None, None, None, None, None, None, None, None)
    if rc != 0:
        raise TpsbError(f"tpsb_mk_partition failed ({rc})")
    ne, nh = sz.num_elems, sz.num_nbr_elems
    ev = np.zeros((ne + nh, 8), np.int32)
    xyz = np.zeros((ne + nh, 8, 3))
    gid = np.zeros(ne + nh, np.int64)
    f = [np.zeros(6 * (ne + nh), np.int32) for _ in range(4)]
    nbr = np.zeros(max(sz.num_nbr_ranks, 1), np.int32)
    so, ro = np.zeros(sz.num_nbr_ranks + 1, np.int32), np.zeros(sz.num_nbr_ranks + 1, np.int32)
    se = np.zeros(max(sz.num_send, 1), np.int32)
    rc = L.tpsb_mk_partition(*args, _ip(ev), _dp(xyz), gid.ctypes.data_as(C.POINTER(C.c_int64)), _ip(f[0]), _ip(f[1]),
                             _ip(f[2]), _ip(f[3]), _ip(nbr), _ip(so), _ip(se), _ip(ro))
    if rc != 0:
        raise TpsbError(f"tpsb_mk_partition failed ({rc})")
    nf = sz.num_faces
    return dict(num_elems=ne, num_nbr_elems=nh, elem_verts=ev, elem_xyz=xyz, elem_gid=gid,
                face_el1=f[0][:nf].copy(), face_el2=f[1][:nf].copy(), face_inf1=f[2][:nf].copy(),
                face_inf2=f[3][:nf].copy(), nbr_rank=nbr[:sz.num_nbr_ranks].copy(), send_offset=so, recv_offset=ro,
                send_elems=se[:sz.num_send].copy())


def partition_elements(mesh, nparts, method="metis"):
    """Element -> rank map of a hexahedral mesh: METIS k-way on the dual graph (what MFEM's GeneratePartitioning calls,
    src/M2ulPhyS.cpp:332) or recursive coordinate bisection.  Returns (elem_rank int32[NE], edge cut or None)."""
    L = lib()
    ne = mesh["elem_xyz"].shape[0]
    rank = np.zeros(ne, np.int32)
    if method == "metis":
        el1 = np.ascontiguousarray(mesh["face_el1"], np.int32)
        el2 = np.ascontiguousarray(mesh["face_el2"], np.int32)
        cut = C.c_int64(0)
        rc = L.tpsb_mk_partition_metis(ne, len(el1), _ip(el1), _ip(el2), nparts, _ip(rank), C.byref(cut))
        if rc != 0:
            raise TpsbError(f"tpsb_mk_partition_metis failed ({rc})")
        return rank, int(cut.value)
    xyz = np.ascontiguousarray(mesh["elem_xyz"], np.float64)
    rc = L.tpsb_mk_partition_rcb_dim(xyz.shape[2], ne, _dp(xyz), nparts, _ip(rank))
    if rc != 0:
        raise TpsbError(f"tpsb_mk_partition_rcb failed ({rc})")
    return rank, None


def partition_mesh(mesh, elem_rank, rank):
    """meshkit: this rank's piece of a global hexahedral mesh under an arbitrary element -> rank map, in ParMesh's local
    numbering with its face-neighbour (halo) tables; same dict as cartesian_hex_partition plus 'face_attr' when the
    global mesh carries boundary attributes."""
    L = lib()
    ev = np.ascontiguousarray(mesh["elem_verts"], np.int32)
    xyz = np.ascontiguousarray(mesh["elem_xyz"], np.float64)
    g1, g2 = np.ascontiguousarray(mesh["face_el1"], np.int32), np.ascontiguousarray(mesh["face_el2"], np.int32)
    er = np.ascontiguousarray(elem_rank, np.int32)
    sz = PartSizes()
    dim = xyz.shape[2]
    head = (dim, ev.shape[0], _ip(ev), _dp(xyz), len(g1), _ip(g1), _ip(g2), _ip(er), rank, C.byref(sz))
    rc = L.tpsb_mk_partition_general_dim(*head, *([None] * 12))
    if rc != 0:
        raise TpsbError(f"tpsb_mk_partition_general failed ({rc})")
    ne, nh, nf = sz.num_elems, sz.num_nbr_elems, sz.num_faces
    lev, lxyz, gid = np.zeros((ne + nh, 1 << dim), np.int32), np.zeros((ne + nh, 1 << dim, dim)), np.zeros(ne + nh, np.int64)
    f = [np.zeros(max(nf, 1), np.int32) for _ in range(5)]
    nbr = np.zeros(max(sz.num_nbr_ranks, 1), np.int32)
    so, ro = np.zeros(sz.num_nbr_ranks + 1, np.int32), np.zeros(sz.num_nbr_ranks + 1, np.int32)
    se = np.zeros(max(sz.num_send, 1), np.int32)
    rc = L.tpsb_mk_partition_general_dim(*head, _ip(lev), _dp(lxyz), gid.ctypes.data_as(C.POINTER(C.c_int64)), _ip(f[0]), _ip(f[1]),
                                     _ip(f[2]), _ip(f[3]), _ip(f[4]), _ip(nbr), _ip(so), _ip(se), _ip(ro))
    if rc != 0:
        raise TpsbError(f"tpsb_mk_partition_general failed ({rc})")
    out = dict(num_elems=ne, num_nbr_elems=nh, elem_verts=lev, elem_xyz=lxyz, elem_gid=gid,
               face_el1=f[0][:nf].copy(), face_el2=f[1][:nf].copy(), face_inf1=f[2][:nf].copy(), face_inf2=f[3][:nf].copy(),
               face_gface=f[4][:nf].copy(), nbr_rank=nbr[:sz.num_nbr_ranks].copy(), send_offset=so, recv_offset=ro,
               send_elems=se[:sz.num_send].copy())
    if dim == 2:
        out["dim"] = 2
    if "face_attr" in mesh:
        out["face_attr"] = np.ascontiguousarray(np.asarray(mesh["face_attr"], np.int32)[out["face_gface"]])
    return out


def make_halo_desc(part, nccl_comm):
    """tpsb_halo_desc over the arrays of cartesian_hex_partition (keeps them alive on the struct)."""
    h = HaloDesc(len(part["nbr_rank"]), _ip(part["nbr_rank"]), _ip(part["send_offset"]), _ip(part["send_elems"]),
                 _ip(part["recv_offset"]), nccl_comm)
    h._keep = part
    return h


class RhsOperator:
    """Python face of the reference's RHSoperator (src/rhs_operator.hpp:146-184) over the C ABI:
    Mult / updatePrimitives / updateGradients / getGradients, on torch CUDA tensors."""

    def __init__(self, mesh, order=3, physics=None, device=0, halo=None, num_nbr_elems=0, stream=None, bcs=None,
                 face_attr=None, use_bc_in_grad=False, basis_type=0, int_rule_type=0, nvel=None):
        import torch
        self.torch = torch
        self.L = lib()
        self.device = device
        self.physics = physics or Physics.dry_air()
        self._keep = [np.ascontiguousarray(mesh["elem_xyz"], dtype=np.float64)] + [
            np.ascontiguousarray(mesh[k], dtype=np.int32) for k in ("face_el1", "face_el2", "face_inf1", "face_inf2")]
        xyz, el1, el2, i1, i2 = self._keep
        self.NE = xyz.shape[0] - num_nbr_elems
        fattr = None
        if face_attr is not None:
            fattr = np.ascontiguousarray(face_attr, dtype=np.int32)
            self._keep.append(fattr)
        self.dim = xyz.shape[2]  # elem_xyz is [NE][2^dim][dim]
        maps = MeshMaps(self.dim, self.NE, num_nbr_elems, _dp(xyz), len(el1), _ip(el1), _ip(el2), _ip(i1), _ip(i2),
                        _ip(fattr) if fattr is not None else None)
        bcset = None
        if bcs:
            arr = (BcDesc * len(bcs))(*bcs)
            bcset = BcSet(len(bcs), arr, int(use_bc_in_grad))
            self._keep.append(arr)
        self.nvel = nvel or self.dim  # nvel = 3 on a 2-D mesh: axisymmetric run
        neq = self.nvel + 2
        if self.physics.fluid == 1:
            pm = self.physics.plasma.contents
            neq += (pm.num_species - 2 if pm.ambipolar else pm.num_species - 1) + (1 if pm.two_temperature else 0)
        space = SpaceDesc(order, basis_type, int_rule_type, neq, self.nvel)
        self.ctx = C.c_void_p()
        s = stream if stream is not None else 0
        rc = self.L.tpsb_create(C.byref(maps), C.byref(space), C.byref(self.physics),
                                C.byref(bcset) if bcset is not None else None,
                                C.byref(halo) if halo is not None else None, device, C.c_void_p(s), C.byref(self.ctx))
        if rc != 0:
            raise TpsbError(f"tpsb_create failed ({rc}): {self.L.tpsb_last_error(None).decode()}")
        self.N = self.L.tpsb_num_dofs(self.ctx)
        self.neq = self.L.tpsb_num_equation(self.ctx)

    def close(self):
        if getattr(self, "ctx", None):
            self.L.tpsb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        self.close()

    def _chk(self, rc, what):
        if rc != 0:
            raise TpsbError(f"{what} failed ({rc}): {self.L.tpsb_last_error(self.ctx).decode()}")

    def Mult(self, x, y=None):
        """RHSoperator::Mult on device tensors (float64, neq*N)."""
        if y is None:
            y = self.torch.empty_like(x)
        self._chk(self.L.tpsb_rhs_mult(self.ctx, x.data_ptr(), y.data_ptr()), "tpsb_rhs_mult")
        return y

    def mult_host(self, h_x, h_y):
        """Same call on host buffers (numpy or pinned torch tensors)."""
        px = h_x.ctypes.data if isinstance(h_x, np.ndarray) else h_x.data_ptr()
        py = h_y.ctypes.data if isinstance(h_y, np.ndarray) else h_y.data_ptr()
        self._chk(self.L.tpsb_rhs_mult_host(self.ctx, px, py), "tpsb_rhs_mult_host")
        return h_y

    def updatePrimitives(self, x):
        self._chk(self.L.tpsb_update_primitives(self.ctx, x.data_ptr()), "tpsb_update_primitives")

    def updateGradients(self, x, primitives_updated=False):
        self._chk(self.L.tpsb_update_gradients(self.ctx, x.data_ptr(), int(primitives_updated)), "tpsb_update_gradients")

    def _view(self, ptr, n):
        import torch
        # wrap context-owned device memory without copying
        arr = (C.c_double * 0).from_address(0)  # placeholder to keep flake quiet
        del arr
        iface = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}
        holder = type("_CudaView", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=f"cuda:{self.device}")

    def fields(self):
        """(Up, gradUp) views of the context-owned primitive and gradient fields."""
        up, g = C.c_void_p(), C.c_void_p()
        self._chk(self.L.tpsb_get_fields(self.ctx, C.byref(up), C.byref(g)), "tpsb_get_fields")
        return self._view(up.value, self.neq * self.N), self._view(g.value, self.dim * self.neq * self.N)

    def set_solution_view(self, U):
        """The solution grid function U_ the forcing terms read (None: the vector passed to Mult)."""
        self._sol = U
        self._chk(self.L.tpsb_set_solution_view(self.ctx, U.data_ptr() if U is not None else None), "tpsb_set_solution_view")

    def set_distance_field(self, dist):
        """Nodal wall distance (device tensor of N doubles, kept alive here) for the mixing-length model; None: zero."""
        self._dist = dist
        self._chk(self.L.tpsb_set_distance_field(self.ctx, dist.data_ptr() if dist is not None else None), "tpsb_set_distance_field")

    def set_time_step(self, dt):
        """BoundaryCondition::dt: the step the non-reflecting inlets / outlets advance their boundary states with."""
        self._chk(self.L.tpsb_set_time_step(self.ctx, float(dt)), "tpsb_set_time_step")

    def bc_state(self, attr):
        """(meanUp[neq], boundaryU[points, neq]) of the non-reflecting condition on boundary attribute attr."""
        n = C.c_int(0)
        mean = np.zeros(self.neq)
        self._chk(self.L.tpsb_get_bc_state(self.ctx, int(attr), _dp(mean), None, 0, C.byref(n)), "tpsb_get_bc_state")
        bu = np.zeros((max(n.value, 1), self.neq))
        self._chk(self.L.tpsb_get_bc_state(self.ctx, int(attr), _dp(mean), _dp(bu), n.value, C.byref(n)), "tpsb_get_bc_state")
        return mean, bu[:n.value]

    def set_reaction_rate_field(self, rates):
        """Chemistry::setGridFunctionRates: device tensor [components][N] (or None)."""
        self._rates = rates
        ncomp = 0 if rates is None else rates.numel() // self.N
        self._chk(self.L.tpsb_set_reaction_rate_field(self.ctx, rates.data_ptr() if rates is not None else None, ncomp),
                  "tpsb_set_reaction_rate_field")

    def add_forcing(self, kind, **kw):
        """Register a forcing term (tpsb_add_forcing): 'pressure_gradient' (g), 'heat_source' (point1, point2, radius, value),
        'joule_heating' (field: device tensor), 'sponge_zone' (normal, point0, point_init, type, mixed_out, r1, r2, tol, mult,
        target = (rho, u, v, w, p))."""
        d = ForcingDesc()
        d.kind = {"pressure_gradient": 0, "heat_source": 1, "joule_heating": 2, "sponge_zone": 3}[kind]

        def put(dst, src):
            for i, v in enumerate(src):
                dst[i] = float(v)
        if kind == "pressure_gradient":
            put(d.pressure_grad, kw["g"])
        elif kind == "heat_source":
            put(d.hs_point1, kw["point1"]), put(d.hs_point2, kw["point2"])
            d.hs_radius, d.hs_value = float(kw["radius"]), float(kw["value"])
        elif kind == "joule_heating":
            self._jh = kw["field"]
            d.joule_heating = self._jh.data_ptr()
        else:
            put(d.sz_normal, kw["normal"]), put(d.sz_point0, kw["point0"]), put(d.sz_point_init, kw["point_init"])
            d.sz_type, d.sz_mixed_out = int(kw.get("type", 0)), int(kw.get("mixed_out", False))
            d.sz_r1, d.sz_r2, d.sz_tol, d.sz_mult = (float(kw.get(k, 0.0)) for k in ("r1", "r2", "tol", "mult"))
            put(d.sz_target, kw.get("target", (0,) * 5))
        self._chk(self.L.tpsb_add_forcing(self.ctx, C.byref(d)), "tpsb_add_forcing")

    def averaging_add_sample(self, mean, vari, ns_mean, ns_vari, inst=None, vari_start=1, vari_components=None, pressure_slot=True):
        """Averaging::addSample on device tensors (mean: fields x N; vari: (co)variances or None); inst None = the context's Up."""
        nf = mean.numel() // self.N
        vc = vari_components if vari_components is not None else (self.physics_nvel if hasattr(self, "physics_nvel") else 3)
        self._chk(self.L.tpsb_averaging_add_sample(self.ctx, inst.data_ptr() if inst is not None else None, nf, mean.data_ptr(),
                                                   vari.data_ptr() if vari is not None else None, vari_start, vc, ns_mean, ns_vari,
                                                   int(pressure_slot)), "tpsb_averaging_add_sample")

    def clear_forcings(self):
        self._chk(self.L.tpsb_clear_forcings(self.ctx), "tpsb_clear_forcings")

    def mean_time_derivatives(self, y):
        """RHSoperator::getLocalTimeDerivatives: mean |dU/dt| per equation."""
        out = (C.c_double * self.neq)()
        self._chk(self.L.tpsb_get_mean_time_derivatives(self.ctx, y.data_ptr(), out), "tpsb_get_mean_time_derivatives")
        return np.array(out[:])

    def point_eval(self, what, U, aux=None):
        """Test hook: per-point physics of this context on device arrays U [n][neq] (aux: gradUp [n][dim*neq])."""
        which = {"prim": 0, "max_char_speed": 1, "conv_flux": 2, "visc_flux": 3, "source": 4}[what]
        n = U.shape[0]
        shape = {0: (n, self.neq), 1: (n,), 2: (n, self.dim * self.neq), 3: (n, self.dim * self.neq), 4: (n, self.neq)}[which]
        out = self.torch.zeros(shape, dtype=self.torch.float64, device=U.device)
        self._chk(self.L.tpsb_debug_point_eval(self.ctx, which, n, U.data_ptr(), aux.data_ptr() if aux is not None else None,
                                               out.data_ptr()), "tpsb_debug_point_eval")
        return out

    def debug_buffer(self, which):
        """Test hook: device view of an internal buffer (0 face residuals, 1 face-trace blocks)."""
        ptr, cnt = C.c_void_p(), C.c_int64(0)
        self._chk(self.L.tpsb_debug_buffer(self.ctx, which, C.byref(ptr), C.byref(cnt)), "tpsb_debug_buffer")
        return self._view(ptr.value, cnt.value) if ptr.value else None

    def max_char_speed(self):
        out = C.c_double(0.0)
        self._chk(self.L.tpsb_get_max_char_speed(self.ctx, C.byref(out)), "tpsb_get_max_char_speed")
        return out.value

    def ode_step(self, U, dt, scheme=4, nsteps=1):
        self._chk(self.L.tpsb_ode_step(self.ctx, U.data_ptr(), dt, scheme, nsteps), "tpsb_ode_step")
        return U

    def hmin(self):
        out = C.c_double(0.0)
        self._chk(self.L.tpsb_get_hmin(self.ctx, C.byref(out)), "tpsb_get_hmin")
        return out.value

    def check_state(self, U):
        """Check_NAN + Check_Undershoot: number of NaN entries; mixtures get negative species densities clamped."""
        nan = C.c_int(0)
        self._chk(self.L.tpsb_check_state(self.ctx, U.data_ptr(), C.byref(nan)), "tpsb_check_state")
        return nan.value

    def solve_step(self, U, dt, scheme=4, cfl=0.0):
        """M2ulPhyS::solveStep without the I/O: (number of NaN entries, next time step)."""
        nan, nxt = C.c_int(0), C.c_double(0.0)
        self._chk(self.L.tpsb_solve_step(self.ctx, U.data_ptr(), dt, scheme, cfl, C.byref(nan), C.byref(nxt)), "tpsb_solve_step")
        return nan.value, nxt.value

    def element_to_faces(self):
        out = np.zeros(7 * self.NE, dtype=np.int32)
        self._chk(self.L.tpsb_get_element_to_faces(self.ctx, _ip(out)), "tpsb_get_element_to_faces")
        return out

    KERNEL_CLASSES = ("prim", "grad", "face_flux", "elem_resid", "pack", "axpy")

    def set_profiling(self, on):
        self._chk(self.L.tpsb_set_profiling(self.ctx, int(on)), "tpsb_set_profiling")

    def kernel_times(self):
        """{class: (total ms, launches)} accumulated since the last call (device timers)."""
        ms = (C.c_double * 6)()
        cnt = (C.c_int64 * 6)()
        self._chk(self.L.tpsb_get_kernel_times(self.ctx, ms, cnt), "tpsb_get_kernel_times")
        return {k: (ms[i], cnt[i]) for i, k in enumerate(self.KERNEL_CLASSES)}

    def launch_count(self):
        return self.L.tpsb_launch_count(self.ctx)

    PATHS = ("general", "fast", "fused", "generic")

    def path(self):
        """Kernel set selected at create: general | fast | fused | generic."""
        return self.PATHS[self.L.tpsb_get_path(self.ctx)]
