"""Molecular integrals for s-type contracted Gaussians.

pyscf / qiskit-nature are not available on the build or GPU boxes, so the reference's own test
system -- H2 at 0.735 Angstrom in 6-31G, `tests/test_optorbvqe.py:27-31`, whose basis holds s
functions only -- is rebuilt here from closed-form integrals: overlap, kinetic, nuclear attraction
and electron repulsion over s Gaussians need nothing beyond the Boys function F0.  A restricted
Hartree-Fock calculation supplies the molecular orbitals, as `PySCFDriver.run()` hands the reference
MO-basis integrals (`tests/test_optorbvqe.py:33-38`) and the initial partial unitary is the first N
MOs (`base_opt_orb_solver.py:93-103`).

Output convention = the reference's: spatial h[p,q] and g[p,q,r,s] = -1/2 (ps|qr)
(`base_opt_orb_solver.py:89-90`: "++--" coefficients in physicist order, times -1), embedded into
spin-orbital tensors with `synthetic.spin_orbital_integrals`.

This is input generation (tiny, numpy, float64), not part of the hot path.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np
import torch

BOHR_PER_ANGSTROM = 1.0 / 0.52917721092          # pyscf's constant (CODATA 2010)

# 6-31G hydrogen: a 3-primitive contracted s shell and one diffuse s primitive
H_631G = (
    ((18.7311370, 0.03349460), (2.8253937, 0.23472695), (0.6401217, 0.81375733)),
    ((0.1612778, 1.0),),
)
H_STO3G = (
    ((3.42525091, 0.15432897), (0.62391373, 0.53532814), (0.16885540, 0.44463454)),
)


def _boys0(t: np.ndarray) -> np.ndarray:
    t = np.asarray(t, dtype=np.float64)
    small = t < 1e-12
    ts = np.where(small, 1.0, t)
    val = 0.5 * np.sqrt(np.pi / ts) * np.vectorize(math.erf)(np.sqrt(ts))
    return np.where(small, 1.0 - t / 3.0, val)


def s_gaussian_integrals(centers: Sequence[Sequence[float]], charges: Sequence[float],
                         shells: Sequence[Tuple[int, Sequence[Tuple[float, float]]]]):
    """AO integrals over contracted s Gaussians.

    centers: nuclear positions (bohr); charges: nuclear charges; shells: (centre index,
    ((exponent, coefficient), ...)) per basis function.  Returns S, T, V ([M,M]), the chemist-order
    repulsion integrals (pq|rs) [M,M,M,M] and the nuclear repulsion energy."""
    centers = np.asarray(centers, dtype=np.float64)
    prim_a, prim_c, prim_R, owner = [], [], [], []
    for f, (ci, prims) in enumerate(shells):
        for a, c in prims:
            prim_a.append(a)
            prim_c.append(c * (2.0 * a / np.pi) ** 0.75)      # normalised primitive
            prim_R.append(centers[ci])
            owner.append(f)
    a = np.array(prim_a); c = np.array(prim_c); R = np.array(prim_R); owner = np.array(owner)
    n, M = len(a), len(shells)
    p = a[:, None] + a[None, :]
    mu = a[:, None] * a[None, :] / p
    R2 = ((R[:, None, :] - R[None, :, :]) ** 2).sum(-1)
    K = np.exp(-mu * R2)                                       # Gaussian product prefactor
    Pc = (a[:, None, None] * R[:, None, :] + a[None, :, None] * R[None, :, :]) / p[..., None]
    S = (np.pi / p) ** 1.5 * K
    T = mu * (3.0 - 2.0 * mu * R2) * S
    V = np.zeros((n, n))
    for Z, C in zip(charges, centers):
        V -= Z * (2.0 * np.pi / p) * K * _boys0(p * ((Pc - C) ** 2).sum(-1))
    PQ2 = ((Pc[:, :, None, None, :] - Pc[None, None, :, :, :]) ** 2).sum(-1)
    pp, qq = p[:, :, None, None], p[None, None, :, :]
    eri = (2.0 * np.pi ** 2.5 / (pp * qq * np.sqrt(pp + qq)) * K[:, :, None, None] *
           K[None, None, :, :] * _boys0(pp * qq / (pp + qq) * PQ2))
    # contract primitives -> basis functions
    Cm = np.zeros((n, M))
    Cm[np.arange(n), owner] = c
    S, T, V = (Cm.T @ X @ Cm for X in (S, T, V))
    eri = np.einsum("abcd,ap,bq,cr,ds->pqrs", eri, Cm, Cm, Cm, Cm, optimize=True)
    # contracted functions are renormalised (the tabulated coefficients are only normalised to ~1e-8)
    d = 1.0 / np.sqrt(np.diag(S))
    S, T, V = (X * d[:, None] * d[None, :] for X in (S, T, V))
    eri = eri * d[:, None, None, None] * d[None, :, None, None] * d[None, None, :, None] * d[None, None, None, :]
    enuc = 0.0
    for i in range(len(charges)):
        for j in range(i):
            enuc += charges[i] * charges[j] / np.linalg.norm(centers[i] - centers[j])
    return S, T, V, eri, enuc


def rhf(S, hcore, eri, nocc: int, tol: float = 1e-13, maxiter: int = 200):
    """Closed-shell Hartree-Fock with symmetric orthogonalisation.  Returns (electronic energy,
    MO coefficients [M,M] sorted by orbital energy, orbital energies)."""
    s, Us = np.linalg.eigh(S)
    X = Us @ np.diag(s ** -0.5) @ Us.T
    F, E_old, C, eps = hcore, 0.0, None, None
    for _ in range(maxiter):
        eps, Cp = np.linalg.eigh(X.T @ F @ X)
        C = X @ Cp
        Dm = 2.0 * C[:, :nocc] @ C[:, :nocc].T
        J = np.einsum("pqrs,rs->pq", eri, Dm)
        Kx = np.einsum("prqs,rs->pq", eri, Dm)
        F = hcore + J - 0.5 * Kx
        E = 0.5 * np.sum(Dm * (hcore + F))
        if abs(E - E_old) < tol:
            break
        E_old = E
    # fix the arbitrary sign of each MO (largest component positive) so fixtures are reproducible
    for k in range(C.shape[1]):
        if C[np.argmax(np.abs(C[:, k])), k] < 0:
            C[:, k] = -C[:, k]
    return float(E), C, eps


def hydrogen_chain(n_atoms: int = 2, spacing_angstrom: float = 0.735, basis=H_631G):
    """MO-basis integrals of a linear H_n chain in an s-only basis, in the reference's conventions.

    Returns dict(h [M,M], g [M,M,M,M] with g[p,q,r,s] = -1/2 (ps|qr), e_nuc, e_hf (electronic),
    n_alpha, n_beta, mo_coeff) as float64 torch tensors / floats.  n_atoms=2, 0.735 Angstrom,
    6-31G is the system of the reference's tests (M = 4)."""
    if n_atoms % 2:
        raise ValueError("closed-shell chains only (even number of hydrogens)")
    centers = [(0.0, 0.0, i * spacing_angstrom * BOHR_PER_ANGSTROM) for i in range(n_atoms)]
    shells = [(i, prims) for i in range(n_atoms) for prims in basis]
    S, T, V, eri, enuc = s_gaussian_integrals(centers, [1.0] * n_atoms, shells)
    e_hf, C, _ = rhf(S, T + V, eri, n_atoms // 2)
    h_mo = C.T @ (T + V) @ C
    eri_mo = np.einsum("abcd,ap,bq,cr,ds->pqrs", eri, C, C, C, C, optimize=True)
    g = -0.5 * eri_mo.transpose(0, 2, 3, 1)                   # g[p,q,r,s] = -1/2 (ps|qr)
    return {
        "h": torch.from_numpy(np.ascontiguousarray(0.5 * (h_mo + h_mo.T))),
        "g": torch.from_numpy(np.ascontiguousarray(g)),
        "e_nuc": float(enuc), "e_hf": e_hf, "n_alpha": n_atoms // 2, "n_beta": n_atoms // 2,
        "mo_coeff": torch.from_numpy(C),
    }
