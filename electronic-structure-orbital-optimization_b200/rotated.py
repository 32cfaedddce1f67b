"""Rotated-Hamiltonian integrals from the device-resident ERI tensor
and the hooks that make the reference use them.

The reference rebuilds the active-space Hamiltonian once per outer iteration with two CPU einsums
over the full (2M)^4 tensor (base_opt_orb_solver.py:597-604; the same pair of einsums initialises
MCVQE, opt_orb_mcvqe.py:90-98):

    h' = W^T h W,   g'_{ijkl} = sum g_pqrs W_pi W_qj W_rk W_sl,   W = block_diag(U, U)

Here h', g' come from `oo_transform` on the engine that already holds the spatial tensor (K1 in tile
mode + q-contraction + one N x M x N^3 contraction; all-reduced when the tensor is sharded) and are
re-embedded into the reference's spin-blocked Q^4 layout on the host for its qiskit tail
(base_opt_orb_solver.py:606-612).

Three levels, pick the lowest that fits:
  rotated_spin_integrals(optimizer, h, g, U)      -> (h' [Q,Q], g' [Q]^4) numpy arrays
  RotatedHamiltonianMixin                         -> mix into a BaseOptOrbSolver subclass
  patch_reference(BaseOptOrbSolver[, OptOrbMCVQE]) -> monkey-patch the loaded reference classes
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import ingest


def expand_spin_blocks(h_rot: torch.Tensor, g_rot: torch.Tensor, structure) -> Tuple[np.ndarray, np.ndarray]:
    """Spatial (h' [N,N], g' [N]^4) -> the reference's spin-orbital layout: h' on the two diagonal
    spin blocks, g' on the spin blocks that were non-zero in the input tensor (`structure.blocks`),
    zero elsewhere -- exactly what the einsum with W = block_diag(U,U) produces."""
    N = h_rot.shape[0]
    Q = 2 * N
    h_rot = h_rot.detach().to("cpu").numpy()
    g_rot = g_rot.detach().to("cpu").numpy()
    hs = np.zeros((Q, Q))
    hs[:N, :N] = h_rot
    hs[N:, N:] = h_rot
    gs = np.zeros((Q, Q, Q, Q))
    for blk in structure.blocks:
        sl = tuple(slice(b * N, (b + 1) * N) for b in blk)
        gs[sl] = g_rot
    return hs, gs


def rotated_spin_integrals(optimizer, one_body_integrals, two_body_integrals,
                           partial_unitary: torch.Tensor) -> Tuple[np.ndarray, np.ndarray]:
    """(h' [Q,Q], g' [Q]^4) as numpy arrays, the tensors of base_opt_orb_solver.py:597-604, through
    the engine `optimizer` (an esoo_b200.PartialUnitaryProjectionOptimizer) caches for these
    integrals; no RDMs are needed.  Collective when the optimiser runs sharded."""
    optimizer._n_active = int(partial_unitary.shape[1])
    eng, structure = optimizer._engine_for(one_body_integrals, two_body_integrals)
    h_rot, g_rot = eng.transform(partial_unitary.detach().to(torch.float64))
    return expand_spin_blocks(h_rot, g_rot, structure)


class RotatedHamiltonianMixin:
    """Mix into a subclass of the reference's BaseOptOrbSolver (or OptOrbVQE / OptOrbSSVQE / ...):

        class FastOptOrbVQE(esoo_b200.RotatedHamiltonianMixin, OptOrbVQE):
            pass

    `get_rotated_hamiltonian` then takes h', g' from the CUDA engine of the solver's
    partial_unitary_optimizer instead of the two CPU einsums; the qiskit tail (ElectronicEnergy,
    normal ordering, mapper) is the reference's own (base_opt_orb_solver.py:606-612)."""

    def _oo_optimizer(self):
        for opt in getattr(self, "_partial_unitary_optimizer_list", []):
            if opt is not None and hasattr(opt, "_engine_for"):
                return opt
        raise RuntimeError("no esoo_b200.PartialUnitaryProjectionOptimizer left in this solver")

    def rotated_integral_tensors(self, partial_unitary: torch.Tensor):
        return rotated_spin_integrals(self._oo_optimizer(), self.one_body_integrals,
                                      self.two_body_integrals, partial_unitary)

    def get_rotated_hamiltonian(self, partial_unitary: torch.Tensor):
        # same statements as base_opt_orb_solver.py:606-612 with the tensors from the GPU
        from qiskit_nature.second_q.hamiltonians import ElectronicEnergy
        h_rot, g_rot = self.rotated_integral_tensors(partial_unitary)
        num_MO = int(self.num_spin_orbitals / 2)
        energy = ElectronicEnergy.from_raw_integrals(
            h1_a=h_rot[0:num_MO, 0:num_MO],
            h2_aa=-2 * g_rot[0:num_MO, 0:num_MO, 0:num_MO, 0:num_MO])
        return self.mapper.map(energy.second_q_op().normal_order())


def patch_reference(base_solver_cls, mcvqe_cls=None) -> None:
    """Monkey-patch the loaded reference classes in place: BaseOptOrbSolver.get_rotated_hamiltonian
    uses the CUDA transform whenever the solver's optimiser is an esoo_b200 one (and falls back to
    the original method otherwise); for OptOrbMCVQE the pre-rotation of the integrals at the end of
    __init__ (opt_orb_mcvqe.py:90-102) is redone the same way."""
    original = base_solver_cls.get_rotated_hamiltonian

    def get_rotated_hamiltonian(self, partial_unitary):
        try:
            RotatedHamiltonianMixin._oo_optimizer(self)
        except RuntimeError:
            return original(self, partial_unitary)
        return RotatedHamiltonianMixin.get_rotated_hamiltonian(self, partial_unitary)

    base_solver_cls.get_rotated_hamiltonian = get_rotated_hamiltonian
    base_solver_cls.rotated_integral_tensors = RotatedHamiltonianMixin.rotated_integral_tensors
    base_solver_cls._oo_optimizer = RotatedHamiltonianMixin._oo_optimizer
    if mcvqe_cls is not None:
        original_init = mcvqe_cls.__init__

        def __init__(self, *args, **kwargs):
            original_init(self, *args, **kwargs)
            try:
                opt = RotatedHamiltonianMixin._oo_optimizer(self)
            except RuntimeError:
                return
            h_rot, g_rot = rotated_spin_integrals(opt, self.one_body_integrals,
                                                  self.two_body_integrals,
                                                  self.initial_partial_unitary)
            for solver in self._excited_states_solver_list:
                solver.one_body_integrals = h_rot
                solver.two_body_integrals = g_rot

        mcvqe_cls.__init__ = __init__


__all__ = ["expand_spin_blocks", "rotated_spin_integrals", "RotatedHamiltonianMixin",
           "patch_reference"]
