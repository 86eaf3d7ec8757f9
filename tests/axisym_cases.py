"""Inputs shared by the boundary-condition / axisymmetric tests of the generic path: a non-periodic quadrilateral box
with the four sides as boundary patches (1 x = lo, 2 x = hi, 3 y = lo, 4 y = hi), smooth dry-air states with two or
three velocity components, and the six-species argon mixture of the reference's test/inputs/plasma.ini (BASELINE
config C4: [Ar+, Ar_m, Ar_r, Ar_p, E, Ar], non-ambipolar, two temperatures, constant transport, no reactions)."""
import numpy as np

import oracle_api
import tps_b200

MW_AR, MW_E = 39.948e-3, 5.4858e-7


def box(n=(5, 4), lo=(0.5, -0.4), hi=(1.7, 0.6), warp=0.0):
    m = tps_b200.cartesian_quad_mesh(*n, lo=lo, hi=hi, periodic=(0, 0))
    m["face_attr"] = tps_b200.quad_box_face_attr(m, lo, hi)
    if warp:  # interior vertices only, so the patches stay on the box sides
        x = m["elem_xyz"]
        L = np.asarray(hi) - np.asarray(lo)
        t = (x - np.asarray(lo)) / L
        bump = np.sin(np.pi * t[..., 0]) * np.sin(np.pi * t[..., 1])
        d = np.stack([np.sin(2 * np.pi * t[..., 1]), np.cos(2 * np.pi * t[..., 0])], axis=-1)
        m["elem_xyz"] = np.ascontiguousarray(x + warp * L * bump[..., None] * d)
    return m


def bcs(kind, nvel=2, species=()):
    """kind -> [(attr, kind, type, data)]: the C4 set (inviscid + isothermal walls, subsonic inlet, pressure outlet)
    or all-wall variants."""
    inlet = (1.25, 12.0, 3.0, 1.5) + tuple(species)
    sets = {
        "c4": [(1, 2, 0, ()), (2, 2, 3, (298.15,)), (3, 0, 2, inlet), (4, 1, 0, (101000.0,))],
        "adiabatic": [(1, 2, 2, ()), (2, 2, 2, ()), (3, 2, 0, ()), (4, 2, 3, (350.0,))],
        "inviscid": [(1, 2, 0, ()), (2, 2, 0, ()), (3, 2, 0, ()), (4, 2, 0, ())],
        "slip": [(1, 2, 1, ()), (2, 2, 1, ()), (3, 2, 1, ()), (4, 2, 0, ())],
        # non-reflecting inlet (SUB_DENS_VEL_NR) and mass-flow outlet (SUB_MF_NR): data[8] = refLength, data[9..11] = tangent1
        "nr": [(1, 2, 0, ()), (2, 2, 3, (298.15,)), (3, 0, 6, (1.25, 3.0, 12.0, 0.0, 0, 0, 0, 0, 0.7, 1.0, 0.0, 0.0)),
               (4, 1, 3, (18.0, 0, 0, 0, 0, 0, 0, 0, 0.7, 1.0, 0.0, 0.0))],
    }
    return sets[kind]


def dry_state(xy, nvel=2, seed=20261018, perturb=0.01):
    """byNODES conserved state [rho, rho u (nvel), rho E] of a smooth dry-air field."""
    x, y = xy[:, 0], xy[:, 1]
    rho = 1.2 + 0.1 * np.sin(2 * x) * np.cos(3 * y)
    vel = [25 * np.sin(2 * x) * np.cos(y) + 8, -15 * np.cos(x) * np.sin(2 * y) + 3, 10 * np.cos(x + y) + 2][:nvel]
    p = 101300 + 800 * (np.cos(2 * x) + np.cos(3 * y))
    ke = 0.5 * rho * sum(v * v for v in vel)
    U = np.concatenate([rho] + [rho * v for v in vel] + [p / 0.4 + ke])
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(U * (1 + perturb * rng.uniform(-1, 1, U.shape)))


def argon6_dict():
    """test/inputs/plasma.ini species block in mixture order; constant transport; the ini's reaction block is
    disabled (number_of_reactions = 0)."""
    names = ["Ar+", "Ar_m", "Ar_r", "Ar_p", "E", "Ar"]
    mw = [MW_AR - MW_E, MW_AR, MW_AR, MW_AR, MW_E, MW_AR]
    ch = [1.0, 0.0, 0.0, 0.0, -1.0, 0.0]
    fe = [1.521e6, 1.114e6, 1.126e6, 1.26e6, 0.0, 0.0]
    diff = [1.3e-3, 2.1e-3, 1.7e-3, 1.9e-3, 3.1e-2, 1.5e-3]
    nu = [2.3e3, 1.1e3, 0.7e3, 1.9e3, 0.9e3, 4.1e3]
    sp = [dict(mw=mw[i], charge=ch[i], formation_energy=fe[i], molar_cv=1.5, diffusivity=diff[i], mt_freq=nu[i])
          for i in range(len(names))]
    return dict(ambipolar=False, two_temperature=True, viscosity=2.2e-3, bulk_viscosity=4.0e-4, thermal_conductivity=0.12,
                electron_thermal_conductivity=0.3, species=sp, reactions=[])


def argon6_primitives(xy, nvel=3, seed=20261018):
    """[rho, u (nvel), T_h, n_sp (5 active, mol/m^3), T_e] smooth + 1 % seeded perturbation."""
    x, y = xy[:, 0], xy[:, 1]
    cols = [1.4 + 0.1 * np.sin(2 * x) * np.cos(3 * y)]
    cols += [20 * np.sin(2 * x) * np.cos(y) + 6, -12 * np.cos(x) * np.sin(2 * y) + 2, 8 * np.cos(x + y) + 1][:nvel]
    cols += [900 + 100 * np.cos(2 * x) * np.cos(y)]
    cols += [0.02 + 0.01 * np.sin(x + 2 * y), 0.05 + 0.02 * np.cos(3 * x), 0.03 + 0.01 * np.sin(2 * y),
             0.04 + 0.015 * np.cos(x - y), 0.02 + 0.01 * np.sin(x + 2 * y)]
    cols += [3000 + 600 * np.sin(x + y)]
    up = np.stack(cols, axis=1)
    rng = np.random.default_rng(seed)
    return up * (1 + 0.01 * rng.uniform(-1, 1, up.shape))


def make_pair(m, order, eq, bt, ir, nvel, bc_kind, use_bc_in_grad, mixture=None, gpu=True, kind=None, mixing_length=None,
              want_oracle=True, device=0):
    """(RhsOperator or None, Oracle) on mesh m with boundary-condition set bc_kind.
    mixing_length = (max-mixing-length, Pr_ratio, bulk-multiplier): flow/useMixingLength (reference back end)."""
    nsp_in = ()
    if mixture is not None:
        nsp_in = (0.02 * (MW_AR - MW_E), 0.05 * MW_AR, 0.03 * MW_AR, 0.04 * MW_AR, 0.02 * MW_E)
    specs = bcs(bc_kind, nvel, nsp_in) if bc_kind else []
    if mixture is not None:
        pm = tps_b200.PlasmaModels.from_dict(mixture)
        phys_o, phys_g = oracle_api.mixture_params(pm, eq), tps_b200.Physics.plasma_mixture(pm, eq)
        neq = nvel + 2 + (pm.num_species - 1) + 1
        kind = "ref"
    else:
        phys_o, phys_g = oracle_api.dry_air_params(eq, 3e4, 0.2), tps_b200.Physics.dry_air(eq, 3e4, 0.2)
        neq = nvel + 2
        kind = kind or "port"
    if mixing_length is not None:
        kind = "ref"
        phys_g.with_mixing_length(*mixing_length)
        phys_o.use_mixing_length = 1
        phys_o.max_mixing_length, phys_o.mixing_length_Prt, phys_o.mixing_length_bulk_mult = mixing_length
    orc = None
    if want_oracle:
        orc = oracle_api.Oracle(order, m["elem_xyz"], m["face_el1"], m["face_el2"], m["face_inf1"], m["face_inf2"],
                                phys=phys_o, kind=kind, basis_type=bt, int_rule=ir, neq=neq, nvel=nvel)
    if specs and orc is not None:
        orc.set_bcs(m["face_attr"], [oracle_api.make_bc(*b) for b in specs], use_bc_in_grad)
    op = None
    if gpu:
        op = tps_b200.RhsOperator(m, order=order, physics=phys_g, basis_type=bt, int_rule_type=ir, nvel=nvel, device=device,
                                  face_attr=m["face_attr"] if specs else None, use_bc_in_grad=use_bc_in_grad,
                                  bcs=[tps_b200.BcDesc.make(*b) for b in specs] if specs else None)
    return op, orc
