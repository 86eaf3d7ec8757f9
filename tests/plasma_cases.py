"""Plasma-model inputs shared by the mixture tests: the ternary argon test mixture of the reference's
test/inputs/mms.ternary_plasma.2d.ini (BASELINE config C3), in MIXTURE order [Ar+, E, Ar]
(electron second to last, background last; src/M2ulPhyS.cpp:2979-3137)."""
import numpy as np

import tps_b200

MW_AR, MW_E = 39.948e-3, 10.0e-3  # "This is not a real electron mass. For test purpose." (the ini)


def ternary_dict(ambipolar=True, two_temperature=True):
    return dict(
        ambipolar=ambipolar, two_temperature=two_temperature, viscosity=1.1, bulk_viscosity=0.3, thermal_conductivity=0.6,
        electron_thermal_conductivity=0.3,
        species=[dict(mw=MW_AR - MW_E, charge=1.0, formation_energy=1.521e4, molar_cv=1.5, diffusivity=1.3, mt_freq=2.3),
                 dict(mw=MW_E, charge=-1.0, formation_energy=0.0, molar_cv=1.5, diffusivity=3.1, mt_freq=0.9),
                 dict(mw=MW_AR, charge=0.0, formation_energy=0.0, molar_cv=1.5, diffusivity=1.9, mt_freq=4.1)],
        reactions=[dict(model=0, A=4.7, b=1.2, E=6.49e4, energy=1.521e4, detailed=True, eqA=1.39, eqB=0.7, eqE=6.197e2,
                        reactants=[0, 1, 1], products=[1, 2, 0])])  # Ar + E <=> Ar.+1 + 2 E


def ternary_models(**kw):
    return tps_b200.PlasmaModels.from_dict(ternary_dict(**kw))


def random_primitives(n, dim=2, seed=3):
    """[rho, u(dim), T_h, n_ion, T_e] within the ranges of test/test_boundary_flux.cpp's random states."""
    rng = np.random.default_rng(seed)
    up = np.zeros((n, dim + 4))
    up[:, 0] = rng.uniform(0.9, 1.4, n)
    up[:, 1:1 + dim] = rng.uniform(-40, 40, (n, dim))
    up[:, 1 + dim] = rng.uniform(280, 700, n)
    up[:, 2 + dim] = rng.uniform(0.05, 1.0, n)
    up[:, 3 + dim] = rng.uniform(320, 3000, n)
    return up


def smooth_primitives(xy, seed=20261018):
    """Smooth periodic primitive field on [-pi, pi]^2 plus the +-1 % seeded perturbation."""
    x, y = xy[:, 0], xy[:, 1]
    up = np.stack([1.2 + 0.1 * np.sin(x) * np.cos(y), 20 * np.sin(x) * np.cos(y) + 5, -20 * np.cos(x) * np.sin(y) + 2,
                   400 + 60 * np.cos(x) * np.cos(2 * y), 0.5 + 0.3 * np.sin(2 * x) * np.sin(y),
                   1500 + 500 * np.sin(x + y)], axis=1)
    rng = np.random.default_rng(seed)
    return up * (1 + 0.01 * rng.uniform(-1, 1, up.shape))


def argon_minimal_dict(third_order=False, multipliers=None, two_temperature=True, ambipolar=True):
    """Ternary argon with transport_model = argon_minimal (GasMinimalTransport: collision-integral transport,
    test/inputs/argonMinimal*.ini) and the physical electron mass."""
    me = 5.48579908782496e-7
    d = ternary_dict(ambipolar=ambipolar, two_temperature=two_temperature)
    d["species"][0]["mw"], d["species"][1]["mw"] = MW_AR - me, me
    d["transport_model"] = "argon_minimal"
    d["third_order_k_electron"] = third_order
    if multipliers:
        d["multipliers"] = multipliers
    return d


def hot_primitives(xy, seed=20261018):
    """Weakly ionised hot argon: T_h ~ 9000 K, T_e ~ 11000 K, n_ion ~ 0.02 mol/m^3, rho ~ 0.05 kg/m^3 (n_Ar ~ 1.3)."""
    x, y = xy[:, 0], xy[:, 1]
    up = np.stack([0.052 + 0.004 * np.sin(x) * np.cos(y), 200 * np.sin(x) * np.cos(y) + 50, -200 * np.cos(x) * np.sin(y) + 20,
                   9000 + 800 * np.cos(x) * np.cos(2 * y), 0.02 + 0.008 * np.sin(2 * x) * np.sin(y),
                   11000 + 1500 * np.sin(x + y)], axis=1)
    rng = np.random.default_rng(seed)
    return up * (1 + 0.01 * rng.uniform(-1, 1, up.shape))
