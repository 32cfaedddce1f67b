"""CPU oracle (numpy) for the orbital-optimisation inner loop.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain numpy, the algorithm of the reference implementation
(JoelHBierman/electronic-structure-orbital-optimization).  It is the checker for the CUDA path and
must never be imported by the product package: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it.

Parity status: the reference's own tests hold no tensor-level golden vectors for this path
(SURVEY.md section 8c).  The oracle is therefore pinned against outputs of the *live* reference
code, loaded from /root/reference by ``oracle/ref_loader.py`` and frozen as fixtures under
``tests/golden/`` by ``tests/golden/make_golden.py``.  The end-to-end numbers hard-coded in the
reference's tests (H2 / 6-31G: tests/test_optorbvqe.py:67, tests/test_optorbmcvqe.py:61) are
reproduced too, with this oracle as the optimiser inside the exact-diagonalisation outer loop
(tests/test_oracle_golden.py::test_oracle_outer_loop_on_molecule).

Reference files restated here (paths under electronic_structure_algorithms/orbital_optimization/):
  base_opt_orb_solver.py:534-582   compute_rotated_energy          -> rotated_energy_spin
  opt_orb_eigensolver.py:149-169   compute_rotated_weighted_energy_sum -> weighted_energy_sum_spin
  partial_unitary_projection_optimizer.py:85-103  autograd gradient -> rotated_energy_grad_spin
  partial_unitary_projection_optimizer.py:70-83   orth             -> orth
  partial_unitary_projection_optimizer.py:129-159 BB update        -> bb_update
  partial_unitary_projection_optimizer.py:161-350 driver loop      -> optimal_rotation
  base_opt_orb_solver.py:597-604   rotated integrals               -> rotated_integrals_spin
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "block_diag2", "rotated_energy_spin", "rotated_energy_grad_spin", "weighted_energy_sum_spin",
    "rotated_energy_spatial", "rotated_energy_grad_spatial", "rotated_integrals_spin",
    "rotated_integrals_spatial", "orth", "bb_update", "optimal_rotation",
]


def block_diag2(U: np.ndarray) -> np.ndarray:
    """W = block_diag(U, U): alpha orbitals first, then beta (base_opt_orb_solver.py:549)."""
    M, N = U.shape
    W = np.zeros((2 * M, 2 * N), dtype=U.dtype)
    W[:M, :N] = U
    W[M:, N:] = U
    return W


def _transform_last3(g: np.ndarray, W: np.ndarray) -> np.ndarray:
    """T3[p,j,k,l] = sum_{qrs} g[p,q,r,s] W[q,j] W[r,k] W[s,l]  (staged, cheapest index last)."""
    t = np.tensordot(g, W, axes=([3], [0]))          # p q r l
    t = np.tensordot(t, W, axes=([2], [0]))          # p q l k
    t = np.tensordot(t, W, axes=([1], [0]))          # p l k j
    return np.ascontiguousarray(t.transpose(0, 3, 2, 1))  # p j k l


def _energy_general(W, D, G, h, g):
    """E = sum h_pq W_pi W_qj D_ij + sum g_pqrs W_pi W_qj W_rk W_sl G_ijkl
    (the two einsums of base_opt_orb_solver.py:554-563)."""
    e1 = np.einsum("pq,pi,qj,ij->", h, W, W, D, optimize=True)
    T3 = _transform_last3(g, W)
    A = np.tensordot(T3, G, axes=([1, 2, 3], [1, 2, 3]))  # p a
    return float(e1 + np.sum(W * A))


def _grad_general(W, D, G, h, g):
    """dE/dW for a general (no symmetry assumed) h, g, D, G: one term per index slot."""
    grad = h @ W @ D.T + h.T @ W @ D
    for slot in range(4):
        gs = np.moveaxis(g, slot, 0)
        Gs = np.moveaxis(G, slot, 0)
        T3 = _transform_last3(np.ascontiguousarray(gs), W)
        grad = grad + np.tensordot(T3, Gs, axes=([1, 2, 3], [1, 2, 3]))
    return grad


# ---------------------------------------------------------------------------------------------
# spin-orbital picture: exactly the reference's call signature (tensors of extent 2M / 2N)
# ---------------------------------------------------------------------------------------------
def rotated_energy_spin(U, oneRDM, twoRDM, one_body_integrals, two_body_integrals) -> float:
    """Real branch of compute_rotated_energy (base_opt_orb_solver.py:549-563)."""
    return _energy_general(block_diag2(U), oneRDM, twoRDM, one_body_integrals, two_body_integrals)


def rotated_energy_grad_spin(U, oneRDM, twoRDM, one_body_integrals, two_body_integrals):
    """What torch.autograd.grad of the energy w.r.t. U returns
    (partial_unitary_projection_optimizer.py:85-103): the gradient flows through block_diag, i.e.
    the alpha-alpha and beta-beta blocks of dE/dW are summed."""
    M, N = U.shape
    gW = _grad_general(block_diag2(U), oneRDM, twoRDM, one_body_integrals, two_body_integrals)
    return gW[:M, :N] + gW[M:, N:]


def weighted_energy_sum_spin(U, oneRDMs, twoRDMs, one_body_integrals, two_body_integrals, weights):
    """compute_rotated_weighted_energy_sum (opt_orb_eigensolver.py:149-169)."""
    total = 0.0
    for w, D, G in zip(weights, oneRDMs, twoRDMs):
        total += float(w) * rotated_energy_spin(U, D, G, one_body_integrals, two_body_integrals)
    return total


def weighted_energy_grad_spin(U, oneRDMs, twoRDMs, one_body_integrals, two_body_integrals, weights):
    total = np.zeros_like(U)
    for w, D, G in zip(weights, oneRDMs, twoRDMs):
        total += float(w) * rotated_energy_grad_spin(U, D, G, one_body_integrals, two_body_integrals)
    return total


def rotated_integrals_spin(U, one_body_integrals, two_body_integrals):
    """h' = W^T h W, g'_{ijkl} = sum g_pqrs W_pi W_qj W_rk W_sl (base_opt_orb_solver.py:597-604)."""
    W = block_diag2(U)
    h_rot = W.T @ one_body_integrals @ W
    T3 = _transform_last3(two_body_integrals, W)
    g_rot = np.tensordot(W, T3, axes=([0], [0]))
    return h_rot, g_rot


# ---------------------------------------------------------------------------------------------
# spatial-orbital picture (what the CUDA path computes; same einsum on M^4 data)
# ---------------------------------------------------------------------------------------------
def rotated_energy_spatial(U, D, G, h, g) -> float:
    return _energy_general(U, D, G, h, g)


def rotated_energy_grad_spatial(U, D, G, h, g):
    return _grad_general(U, D, G, h, g)


def rotated_integrals_spatial(U, h, g):
    h_rot = U.T @ h @ U
    T3 = _transform_last3(g, U)
    return h_rot, np.tensordot(U, T3, axes=([0], [0]))


# ---------------------------------------------------------------------------------------------
# retraction, BB step, driver loop
# ---------------------------------------------------------------------------------------------
def orth(V: np.ndarray) -> np.ndarray:
    """orth(V) = V Q diag(L)^(-1/2) Q^T with (L, Q) = eigh(V^T V)
    (partial_unitary_projection_optimizer.py:80-81)."""
    L, Q = np.linalg.eigh(V.T @ V)
    return V @ Q @ np.diag(1.0 / np.sqrt(L)) @ Q.T


def bb_update(iteration_number, U_cur, U_prev, G_cur, G_prev, stepsize):
    """compute_updated_partial_unitary (partial_unitary_projection_optimizer.py:129-159).
    Returns (U_next, stepsize)."""
    if iteration_number % 2 != 0:
        dU, dG = U_cur - U_prev, G_cur - G_prev
        stepsize = np.trace(dU.T @ dU) / abs(np.trace(dU.T @ dG))
    if iteration_number % 2 == 0 and iteration_number != 0:
        dU, dG = U_cur - U_prev, G_cur - G_prev
        stepsize = abs(np.trace(dU.T @ dG)) / np.trace(dG.T @ dG)
    return orth(U_cur - stepsize * G_cur), stepsize


def optimal_rotation(energy_fn, grad_fn, U0, initial_BBstepsize, stopping_tolerance, maxiter,
                     decay_factor=0.8, callback=None):
    """compute_optimal_rotation (partial_unitary_projection_optimizer.py:161-350).

    Returns a dict with the reference's return values (U, energy) plus bookkeeping:
    'U', 'energy' (= P4_array[0]), 'n_iter' (final iteration_number), 'callbacks' (list of the
    (iteration, energy) pairs a callback would have seen), 'stepsize'.
    """
    tol, d = stopping_tolerance, decay_factor
    P4 = [None, None, None]
    St = [None, 1.5 * tol]
    calls = []

    def cb(k, e):
        calls.append((k, float(e)))
        if callback is not None:
            callback(k, float(e))

    step = initial_BBstepsize
    U_cur, U_prev, G_cur, G_prev = np.array(U0, dtype=np.float64), None, None, None
    k = 0
    P4[2] = energy_fn(U_cur)
    cb(k, P4[2])
    G_cur = grad_fn(U_cur)
    U_new, step = bb_update(k, U_cur, U_prev, G_cur, G_prev, step)
    G_new = grad_fn(U_new)
    U_prev, G_prev, U_cur, G_cur = U_cur, G_cur, U_new, G_new
    k += 1
    P4[1] = energy_fn(U_cur)
    cb(k, P4[1])
    St[0] = (1 - d) * abs(P4[1] - P4[2]) + d * St[1]
    U_new, step = bb_update(k, U_cur, U_prev, G_cur, G_prev, step)
    G_new = grad_fn(U_new)
    U_prev, G_prev, U_cur, G_cur = U_cur, G_cur, U_new, G_new
    k += 1
    P4[0] = energy_fn(U_cur)
    cb(k, P4[0])
    St = [St[1], St[0]]
    St[0] = (1 - d) * abs(P4[0] - P4[1]) + d * St[1]
    U_new, step = bb_update(k, U_cur, U_prev, G_cur, G_prev, step)
    G_new = grad_fn(U_new)
    U_prev, G_prev, U_cur, G_cur = U_cur, G_cur, U_new, G_new
    k += 1
    while St[0] > tol and k <= maxiter:
        P4 = [P4[2], P4[0], P4[1]]          # np.roll(P4, 1)
        P4[0] = energy_fn(U_cur)
        cb(k, P4[1])
        St = [St[1], St[0]]
        St[0] = (1 - d) * abs(P4[1] - P4[2]) + d * St[1]
        U_new, step = bb_update(k, U_cur, U_prev, G_cur, G_prev, step)
        G_new = grad_fn(U_new)
        U_prev, G_prev, U_cur, G_cur = U_cur, G_cur, U_new, G_new
        k += 1
    return {"U": U_cur, "energy": float(P4[0]), "n_iter": k, "callbacks": calls,
            "stepsize": float(step)}
