"""qiskit-free stand-in for the reference's outer loop.

The reference alternates a quantum eigensolver (VQE + Aer estimator) with the orbital optimiser
(opt_orb_minimum_eigensolver.py:150-246, opt_orb_eigensolver.py:171-269).  qiskit / pyscf are not
available here, so the eigensolver is replaced by an exact diagonalisation in the N-orbital active
space; everything around it follows the reference:

    H(U) from the rotated integrals            base_opt_orb_solver.py:584-612
    eigenstate(s) -> 1-/2-RDMs                 base_opt_orb_solver.py:362-532
        D[p,q] = <a+_p a_q>,  G[p,q,r,s] = <a+_p a+_q a_s a_r>   (spin-orbital, alpha block first)
    optimizer.compute_optimal_rotation(fun=..., oneRDM=..., twoRDM=..., one_body_integrals=...,
        two_body_integrals=..., initial_partial_unitary=...)[0]      opt_orb_minimum_eigensolver.py:223-228
    stopping rule on the outer energies        opt_orb_minimum_eigensolver.py:125-138

Any object with the reference's optimiser API can be plugged in: the reference class itself (CPU)
or esoo_b200.PartialUnitaryProjectionOptimizer (CUDA).  The eigensolver side is plain numpy and
tiny (a few dozen determinants): it is plumbing around the hot path, not part of it.
"""
from __future__ import annotations

import copy
import itertools
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------
# second quantisation in a small Fock sector
# ------------------------------------------------------------------------------------------------
class FockSector:
    """All determinants with `nelec` electrons in `Q` spin orbitals (bit strings), with the
    excitation operators E_pq = a+_p a_q as dense matrices."""

    def __init__(self, Q: int, nelec: int):
        self.Q, self.nelec = Q, nelec
        self.dets = [sum(1 << i for i in occ) for occ in itertools.combinations(range(Q), nelec)]
        self.index = {d: i for i, d in enumerate(self.dets)}
        n = len(self.dets)
        self.E = np.zeros((Q, Q, n, n))
        for col, d in enumerate(self.dets):
            for q in range(Q):
                if not (d >> q) & 1:
                    continue
                s1 = (-1) ** bin(d & ((1 << q) - 1)).count("1")
                d1 = d & ~(1 << q)
                for p in range(Q):
                    if (d1 >> p) & 1:
                        continue
                    s2 = (-1) ** bin(d1 & ((1 << p) - 1)).count("1")
                    self.E[p, q, self.index[d1 | (1 << p)], col] = s1 * s2

    def sz_subspace(self, n_alpha: int, N: int) -> np.ndarray:
        """Indices of determinants with n_alpha electrons in spin orbitals [0, N)."""
        mask = (1 << N) - 1
        return np.array([i for i, d in enumerate(self.dets)
                         if bin(d & mask).count("1") == n_alpha], dtype=int)

    def hamiltonian(self, h: np.ndarray, g: np.ndarray) -> np.ndarray:
        """H = sum h_pq a+_p a_q + sum g_pqrs a+_p a+_q a_s a_r,
        using a+_p a+_q a_s a_r = E_pr E_qs - delta_qr E_ps."""
        Q, n = self.Q, len(self.dets)
        Em = self.E.reshape(Q * Q, n, n)
        H = np.tensordot(h.reshape(-1), Em, axes=([0], [0]))
        # sum_pqrs g_pqrs E_pr E_qs
        gperm = g.transpose(0, 2, 1, 3).reshape(Q * Q, Q * Q)       # [(p r), (q s)]
        left = np.tensordot(gperm, Em, axes=([1], [0]))            # [(p r), n, n] = sum_qs g E_qs
        H = H + np.einsum("aij,ajk->ik", Em, left)
        # - sum_pqs g_pqqs E_ps
        gc = np.einsum("pqqs->ps", g)
        H = H - np.tensordot(gc.reshape(-1), Em, axes=([0], [0]))
        return 0.5 * (H + H.T)

    def rdms(self, psi: np.ndarray):
        """D[p,q] = <a+_p a_q>,  G[p,q,r,s] = <a+_p a+_q a_s a_r> = <E_pr E_qs> - delta_qr D_ps."""
        Q = self.Q
        w = np.tensordot(self.E, psi, axes=([3], [0]))            # w[a,b] = E_ab psi
        D = np.tensordot(w, psi, axes=([2], [0]))                 # <psi|E_pq|psi>
        # <E_pr E_qs> = (E_rp psi).(E_qs psi)
        gram = np.tensordot(w, w, axes=([2], [2]))                # [r,p,q,s]
        G = gram.transpose(1, 2, 0, 3).copy()                     # [p,q,r,s]
        for q in range(Q):
            G[:, q, q, :] -= D
        return D, G


def _rotated_spin_integrals(h: torch.Tensor, g: torch.Tensor, U: torch.Tensor, engine=None):
    """Spin-orbital (h' [Q,Q], g' [Q,Q,Q,Q]) for W = block_diag(U,U)
    (base_opt_orb_solver.py:597-604).  With an OrbitalEngine the spatial transform runs on the GPU
    (oo_transform) and is re-embedded into the reference's spin-block layout."""
    M, N = U.shape
    if engine is not None:
        from . import ingest, rotated
        if hasattr(engine, "_engine_for"):          # an esoo_b200 optimiser: its cached engine
            return rotated.rotated_spin_integrals(engine, h, g, U)
        h_rot, g_rot = engine.transform(U)
        _, _, st = ingest.reduce_integrals(h, g)
        return rotated.expand_spin_blocks(h_rot, g_rot, st)
    W = torch.block_diag(U, U)
    h_rot = torch.einsum('pq,pi,qj->ij', h, W, W)
    g_rot = torch.einsum('pqrs,pi,qj,rk,sl->ijkl', g, W, W, W, W)
    return h_rot.numpy(), g_rot.numpy()


class _Solver:
    """Carries what the optimiser reads from the bound objective: the method name and, for the
    state-averaged case, `weight_vector` (opt_orb_eigensolver.py:87-95)."""

    def __init__(self, energy_impl: Optional[Callable], weights=None):
        self._impl = energy_impl
        self.wavefunction_real = True
        if weights is not None:
            self.weight_vector = list(weights)

    def compute_rotated_energy(self, partial_unitary, oneRDM, twoRDM, one_body_integrals,
                               two_body_integrals):
        return self._impl(partial_unitary, oneRDM, twoRDM, one_body_integrals, two_body_integrals)

    def compute_rotated_weighted_energy_sum(self, partial_unitary, oneRDM, twoRDM,
                                            one_body_integrals, two_body_integrals):
        total = 0
        for idx, (d, g2) in enumerate(zip(oneRDM, twoRDM)):
            total = total + torch.tensor(self.weight_vector[idx], dtype=torch.float64) * \
                self._impl(partial_unitary, d, g2, one_body_integrals, two_body_integrals)
        return total


def run_outer_loop(optimizer, h: torch.Tensor, g: torch.Tensor, num_spin_orbitals: int,
                   n_alpha: int, n_beta: int, maxiter: int = 10, stopping_tolerance: float = 1e-5,
                   n_states: int = 1, weights: Optional[Sequence[float]] = None,
                   initial_partial_unitary: Optional[torch.Tensor] = None,
                   energy_impl: Optional[Callable] = None, engine_for_transform=None,
                   outer_loop_callback: Optional[Callable] = None,
                   state_indices: Optional[Sequence[int]] = None):
    """Exact-diagonalisation version of OptOrbMinimumEigensolver.compute_minimum_energy (n_states=1)
    / OptOrbEigensolver.compute_energies (n_states>1, state-averaged with `weights`).

    h, g: the reference's spin-orbital integral tensors ([2M,2M], [2M]^4, float64, CPU).
    energy_impl: the Python objective handed to optimisers that *call* `fun` (the reference class);
    the CUDA optimiser never calls it.  state_indices picks which eigenvectors of the (S_z-resolved)
    active-space Hamiltonian play the role of the k states (default: the k lowest); e.g. (0, 2)
    skips the S_z = 0 triplet component that a singlet-preserving ansatz cannot reach.
    Returns dict(energies=[per outer iteration: list of state
    energies], U=final partial unitary, inner_iterations=[...])."""
    P = h.shape[0]
    M, Q = P // 2, num_spin_orbitals
    N = Q // 2
    if initial_partial_unitary is None:                 # base_opt_orb_solver.py:93-103
        U = torch.zeros(M, N, dtype=torch.float64)
        for n in range(N):
            U[n, n] = 1.0
    else:
        U = initial_partial_unitary.clone()
    if weights is None and n_states > 1:
        weights = [n_states - n for n in range(n_states)]    # opt_orb_eigensolver.py:93
    solver = _Solver(energy_impl, weights if n_states > 1 else None)
    fun = solver.compute_rotated_energy if n_states == 1 else \
        solver.compute_rotated_weighted_energy_sum
    sector = FockSector(Q, n_alpha + n_beta)
    sub = sector.sz_subspace(n_alpha, N)
    optimizers = [copy.deepcopy(optimizer) for _ in range(int(maxiter) + 1)]   # base.py:75
    energies: List[List[float]] = []
    it = 0

    def stop(iteration):                                  # opt_orb_minimum_eigensolver.py:125-138
        if len(energies) >= 2:
            last, prev = energies[-1], energies[-2]
            crit = abs(last[0] - prev[0]) if n_states == 1 else \
                abs(sum(w * e for w, e in zip(weights, last)) -
                    sum(w * e for w, e in zip(weights, prev)))
            return iteration == maxiter or crit < stopping_tolerance
        return False

    while not stop(it):
        h_rot, g_rot = _rotated_spin_integrals(h, g, U, engine_for_transform)
        H = sector.hamiltonian(h_rot, g_rot)
        evals, evecs = np.linalg.eigh(H[np.ix_(sub, sub)])
        states = []
        picks = list(state_indices) if state_indices is not None else list(range(n_states))
        for n in picks:
            psi = np.zeros(len(sector.dets))
            psi[sub] = evecs[:, n]
            states.append(psi)
        energies.append([float(evals[n]) for n in picks])
        if outer_loop_callback is not None:
            outer_loop_callback(it, energies[-1], U)
        if stop(it):
            break
        rd = [sector.rdms(psi) for psi in states]
        Ds = [torch.from_numpy(np.ascontiguousarray(d)) for d, _ in rd]
        Gs = [torch.from_numpy(np.ascontiguousarray(g2)) for _, g2 in rd]
        one = Ds[0] if n_states == 1 else Ds
        two = Gs[0] if n_states == 1 else Gs
        dev = optimizers[it].device
        to = (lambda t: [x.to(dev) for x in t] if isinstance(t, list) else t.to(dev))
        U = optimizers[it].compute_optimal_rotation(
            fun=fun, oneRDM=to(one), twoRDM=to(two), one_body_integrals=h.to(dev),
            two_body_integrals=g.to(dev), initial_partial_unitary=U)[0]
        U = U.detach().to('cpu')
        optimizers[it] = None
        it += 1
    return {"energies": energies, "U": U, "outer_iterations": len(energies)}
