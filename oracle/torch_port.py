"""CPU baseline: the reference's own torch formulation of the path, restated.  TEST/BENCH
INFRASTRUCTURE ONLY (imported by bench.py's cpu_baseline / --impl reference legs and by tests).

The reference evaluates the energy with one 6-operand torch.einsum contracted left to right
(base_opt_orb_solver.py:558-563; opt_einsum is absent, so the order is the written one: p, then q,
r, s) and obtains dE/dU by torch.autograd.grad through it
(partial_unitary_projection_optimizer.py:85-103).  This file states the same computation for the
spatial-orbital tensors the CUDA path consumes.

Bounded sample: every pairwise contraction of that chain is linear in the extent of the LAST index
s, so a slab g[:, :, :, s0:s0+ms] (with the matching rows of the last U operand) costs ms/M of a
full evaluation in every step of the chain; that is how a bounded sample of the M=256 workload is
timed on the host.  (A slab of the first index would not do: the second contraction of the chain
does not shrink with it.)
"""
from __future__ import annotations

import time

import torch


def energy_last_slab(U, D, G, h, g_slab, s0: int):
    """Energy contribution of the slab s in [s0, s0+ms) of g's last index (all of it when the slab
    is the whole tensor), einsum strings as in the reference."""
    ms = g_slab.shape[3]
    Us = U[s0:s0 + ms]
    e1 = torch.einsum('pq,pi,qj,ij', h[:, s0:s0 + ms], U, Us, D)
    e2 = torch.einsum('pqrs,pi,qj,rk,sl,ijkl', g_slab, U, U, U, Us, G)
    return e1 + e2


def energy_and_autograd(U, D, G, h, g_slab, s0: int = 0):
    """(E, dE/dU) the way the reference gets them: forward einsum + torch.autograd.grad."""
    U = U.clone().requires_grad_(True)
    E = energy_last_slab(U, D, G, h, g_slab, s0)
    (grad,) = torch.autograd.grad([E], inputs=[U])
    return float(E.detach()), grad


def time_reference(U, D, G, h, g_slab, s0: int = 0, repeats: int = 1):
    """(seconds per (E, dE/dU) evaluation = forward + backward, seconds per reference optimiser
    iteration = energy-only forward (pupo.py:310) + forward + backward (pupo.py:331))."""
    t_eval = t_iter = 0.0
    for _ in range(repeats):
        a = time.perf_counter()
        with torch.no_grad():
            energy_last_slab(U, D, G, h, g_slab, s0)
        b = time.perf_counter()
        energy_and_autograd(U, D, G, h, g_slab, s0)
        c = time.perf_counter()
        t_eval += c - b
        t_iter += c - a
    return t_eval / repeats, t_iter / repeats


# ------------------------------------------------------------------------------------------------
# The reference's formulation on its own (spin-orbital) tensors: W = block_diag(U, U)
# (base_opt_orb_solver.py:549), the two einsum strings of :554-563, the state-weighted sum of
# opt_orb_eigensolver.py:156-169 and autograd through all of it.  Used to time the reference's
# real problem sizes (BASELINE.json configs 1-3) on the host cores.
# ------------------------------------------------------------------------------------------------
def energy_spin(U, oneRDMs, twoRDMs, weights, h_spin, g_spin):
    W = torch.block_diag(U, U)
    total = 0
    for w, D, G in zip(weights, oneRDMs, twoRDMs):
        e = torch.einsum('pq,pi,qj,ij', h_spin, W, W, D)
        e = e + torch.einsum('pqrs,pi,qj,rk,sl,ijkl', g_spin, W, W, W, W, G)
        total = total + float(w) * e
    return total


def time_reference_spin(U, oneRDMs, twoRDMs, weights, h_spin, g_spin, repeats: int = 1):
    """(seconds per (E, dE/dU), seconds per reference optimiser iteration, E, grad) for the
    spin-orbital problem, every state evaluated in turn as the reference does."""
    t_eval = t_iter = 0.0
    E = grad = None
    for _ in range(repeats):
        a = time.perf_counter()
        with torch.no_grad():
            energy_spin(U, oneRDMs, twoRDMs, weights, h_spin, g_spin)      # pupo.py:310
        b = time.perf_counter()
        Ur = U.clone().requires_grad_(True)
        E = energy_spin(Ur, oneRDMs, twoRDMs, weights, h_spin, g_spin)     # pupo.py:85-103
        (grad,) = torch.autograd.grad([E], inputs=[Ur])
        c = time.perf_counter()
        t_eval += c - b
        t_iter += c - a
    return t_eval / repeats, t_iter / repeats, float(E.detach()), grad
