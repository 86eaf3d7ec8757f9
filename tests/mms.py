"""Manufactured smooth periodic state and its exact Navier-Stokes/Euler right-hand side
(-div(F_c - F_v)) generated with sympy -- the role MASA plays for the reference
(src/masa_handler.cpp, utils/compute_rhs.cpp:104-163); MASA itself is absent here."""
import functools

import numpy as np
import sympy as sp


@functools.lru_cache(maxsize=None)
def _build(eq_system, gamma, R, visc_mult, bulk_mult, C1, S0, Pr):
    x, y, z = sp.symbols("x y z", real=True)
    rho = 1.2 + sp.Rational(1, 10) * sp.sin(x) * sp.cos(y) + sp.Rational(1, 20) * sp.cos(z)
    u = 30 * sp.sin(x) * sp.cos(y) * sp.cos(z) + 10
    v = -30 * sp.cos(x) * sp.sin(y) * sp.cos(z) + 4
    w = 5 * sp.sin(z) * sp.cos(x) - 7
    p = 101300 + 500 * (sp.cos(2 * x) + sp.cos(2 * y)) * (sp.cos(2 * z) + 2)
    vel = [u, v, w]
    X = [x, y, z]
    E = p / (gamma - 1) + rho * (u * u + v * v + w * w) / 2
    U = [rho, rho * u, rho * v, rho * w, E]
    T = p / (rho * R)
    F = [[rho * vel[d] for d in range(3)]]
    for i in range(3):
        F.append([rho * vel[i] * vel[d] + (p if i == d else 0) for d in range(3)])
    F.append([vel[d] * (E + p) for d in range(3)])
    if eq_system != 0:
        mu = C1 * visc_mult * T ** sp.Rational(3, 2) / (T + S0)
        lam = bulk_mult * mu - sp.Rational(2, 3) * mu
        k = gamma * R / (Pr * (gamma - 1)) * mu
        divv = sum(sp.diff(vel[i], X[i]) for i in range(3))
        tau = [[mu * (sp.diff(vel[i], X[j]) + sp.diff(vel[j], X[i])) + (lam * divv if i == j else 0)
                for j in range(3)] for i in range(3)]
        for i in range(3):
            for d in range(3):
                F[1 + i][d] = F[1 + i][d] - tau[i][d]
        for d in range(3):
            F[4][d] = F[4][d] - sum(tau[d][j] * vel[j] for j in range(3)) - k * sp.diff(T, X[d])
    rhs = [-sum(sp.diff(F[e][d], X[d]) for d in range(3)) for e in range(5)]
    prim = [rho, u, v, w, T]
    grad = [[sp.diff(prim[e], X[d]) for d in range(3)] for e in range(5)]
    mods = "numpy"
    fU = sp.lambdify((x, y, z), U, mods, cse=True)
    fR = sp.lambdify((x, y, z), rhs, mods, cse=True)
    fG = sp.lambdify((x, y, z), grad, mods, cse=True)
    return fU, fR, fG


def manufactured(xyz, phys):
    """Returns (U, exact_rhs, exact_gradUp) in the reference's byNODES layout for node coords xyz[N,3]."""
    fU, fR, fG = _build(phys.eq_system, sp.Float(phys.gamma), sp.Float(phys.R), sp.Float(phys.visc_mult),
                        sp.Float(phys.bulk_visc_mult), sp.Float(phys.C1), sp.Float(phys.S0), sp.Float(phys.Pr))
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    one = np.ones_like(x)
    U = np.concatenate([np.asarray(a) * one for a in fU(x, y, z)])
    Rr = np.concatenate([np.asarray(a) * one for a in fR(x, y, z)])
    g = fG(x, y, z)
    G = np.concatenate([np.asarray(g[e][d]) * one for d in range(3) for e in range(5)])
    return U, Rr, G


@functools.lru_cache(maxsize=None)
def _build2d(eq_system, gamma, R, visc_mult, bulk_mult, C1, S0, Pr):
    """2-D counterpart (the reference's mms.euler_2d / ad_cns_2d_sutherlands cases use MASA; regenerated here)."""
    x, y = sp.symbols("x y", real=True)
    rho = 1.2 + sp.Rational(1, 10) * sp.sin(x) * sp.cos(y)
    u = 30 * sp.sin(x) * sp.cos(y) + 10
    v = -30 * sp.cos(x) * sp.sin(y) + 4
    p = 101300 + 500 * (sp.cos(2 * x) + sp.cos(2 * y))
    vel, X = [u, v], [x, y]
    E = p / (gamma - 1) + rho * (u * u + v * v) / 2
    U = [rho, rho * u, rho * v, E]
    T = p / (rho * R)
    F = [[rho * vel[d] for d in range(2)]]
    for i in range(2):
        F.append([rho * vel[i] * vel[d] + (p if i == d else 0) for d in range(2)])
    F.append([vel[d] * (E + p) for d in range(2)])
    if eq_system != 0:
        mu = C1 * visc_mult * T ** sp.Rational(3, 2) / (T + S0)
        lam = bulk_mult * mu - sp.Rational(2, 3) * mu
        k = gamma * R / (Pr * (gamma - 1)) * mu
        divv = sum(sp.diff(vel[i], X[i]) for i in range(2))
        tau = [[mu * (sp.diff(vel[i], X[j]) + sp.diff(vel[j], X[i])) + (lam * divv if i == j else 0)
                for j in range(2)] for i in range(2)]
        for i in range(2):
            for d in range(2):
                F[1 + i][d] = F[1 + i][d] - tau[i][d]
        for d in range(2):
            F[3][d] = F[3][d] - sum(tau[d][j] * vel[j] for j in range(2)) - k * sp.diff(T, X[d])
    rhs = [-sum(sp.diff(F[e][d], X[d]) for d in range(2)) for e in range(4)]
    prim = [rho, u, v, T]
    grad = [[sp.diff(prim[e], X[d]) for d in range(2)] for e in range(4)]
    return (sp.lambdify((x, y), U, "numpy", cse=True), sp.lambdify((x, y), rhs, "numpy", cse=True),
            sp.lambdify((x, y), grad, "numpy", cse=True))


def manufactured2d(xy, phys):
    fU, fR, fG = _build2d(phys.eq_system, sp.Float(phys.gamma), sp.Float(phys.R), sp.Float(phys.visc_mult),
                          sp.Float(phys.bulk_visc_mult), sp.Float(phys.C1), sp.Float(phys.S0), sp.Float(phys.Pr))
    x, y = xy[:, 0], xy[:, 1]
    one = np.ones_like(x)
    U = np.concatenate([np.asarray(a) * one for a in fU(x, y)])
    Rr = np.concatenate([np.asarray(a) * one for a in fR(x, y)])
    g = fG(x, y)
    G = np.concatenate([np.asarray(g[e][d]) * one for d in range(2) for e in range(4)])
    return U, Rr, G
