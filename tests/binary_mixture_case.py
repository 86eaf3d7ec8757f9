"""The reference's end-to-end analytic benchmark test/argon_minimal.binary.test re-expressed:
Ar / Ar+ inter-diffusion wave advected on the doubly periodic 15 x 3 beam mesh (utils/beam_mesh.cpp -rs 0: nt*nx = 15 by
nt = 3 quadrilaterals on 5 x 1), order 3, Gauss-Lobatto basis and rule, argon_minimal transport with the third-order
electron conductivity, 1000 RK4 steps of 6e-5 s (test/inputs/argonMinimal.binary_mixture.ini).  The reference compares
/solution/rho-Y_Ar.+1 with the advected, exponentially damped wave of utils/binary_mixture_ic.cpp:88-152 to a relative
2e-4 (h5diff --relative); the damping rate uses D_ia from MolecularTransport::computeMixtureAverageDiffusivity -- taken
here from the reference's own object code (oracle/_ref)."""
import numpy as np

import tps_b200

RGAS = 8.3144598           # UNIVERSALGASCONSTANT (src/dataStructures.hpp)
MW_AR, MW_E = 39.948e-3, 1.0e-16   # the ini's atoms: "This is not a real electron mass. For test purpose."
P0, T0, U0, LX, LY, KX, KY = 1.0133e0, 300.0, 1.0, 5.0, 1.0, 2.0, 0.0
DT, NSTEPS, TOL = 6.0e-5, 1000, 2e-4


def models_dict():
    sp = [dict(mw=MW_AR - MW_E, charge=1.0, formation_energy=0.0, molar_cv=1.5),   # Ar.+1 = {Ar: 1, E: -1}
          dict(mw=MW_E, charge=-1.0, formation_energy=0.0, molar_cv=1.5),          # E
          dict(mw=MW_AR, charge=0.0, formation_energy=0.0, molar_cv=1.5)]          # Ar (background)
    return dict(ambipolar=False, two_temperature=False, species=sp, reactions=[], transport_model="argon_minimal",
                third_order_k_electron=True)


def mesh():
    return tps_b200.cartesian_quad_mesh(15, 3, lo=(0.0, 0.0), hi=(LX, LY))


def state(xy, decay=1.0, shift=0.0):
    """binary_mixture_ic.cpp:88-117 (initial condition) / :133-152 (reference solution): [rho, rho u, rho v, rho E,
    rho Y_Ar+, rho Y_E]."""
    n_total = P0 / RGAS / T0
    rho = n_total * MW_AR
    rhoU = rho * U0
    rhoE = n_total * (1.5 * RGAS) * T0 + 0.5 * rhoU * rhoU / rho
    Y = 0.5 + 0.45 * decay * np.cos(2.0 * np.pi * KX * (xy[:, 0] - shift) / LX) * np.cos(2.0 * np.pi * KY * xy[:, 1] / LY)
    one = np.ones_like(Y)
    return np.ascontiguousarray(np.concatenate([rho * one, rhoU * one, 0.0 * one, rhoE * one, rho * Y, 0.0 * one]))


def analytic(xy, D_ia, time=DT * NSTEPS):
    decay = np.exp(-(4.0 * np.pi ** 2 * (KX * KX / LX / LX + KY * KY / LY / LY)) * D_ia * time)
    return state(xy, decay, U0 * time), decay
