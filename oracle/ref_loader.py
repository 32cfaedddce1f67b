"""Load the LIVE reference implementation of the hot path from /root/reference.  TEST INFRASTRUCTURE.

Only usable in the build container (the GPU box has no /root/reference): it is used by
``tests/golden/make_golden.py`` to freeze golden vectors and by the CPU-side tests that are skipped
when the reference tree is absent.

``partial_unitary_projection_optimizer.py`` imports only numpy/torch and is loaded by file path.
``base_opt_orb_solver.py`` imports qiskit at module top; the hot-path functions
(``compute_rotated_energy``, ``orth``) use only torch/numpy, so the qiskit modules are replaced by
``MagicMock`` entries in ``sys.modules`` while the file is executed.  The package ``__init__`` is
never imported (it pulls in qiskit).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from unittest import mock

REFERENCE_ROOT = os.environ.get("OO_REFERENCE_ROOT", "/root/reference")
_OO_DIR = os.path.join(REFERENCE_ROOT, "electronic_structure_algorithms", "orbital_optimization")

_STUBS = [
    "qiskit", "qiskit.primitives", "qiskit.quantum_info", "qiskit_algorithms",
    "qiskit_algorithms.variational_algorithm", "qiskit_nature", "qiskit_nature.second_q",
    "qiskit_nature.second_q.mappers", "qiskit_nature.second_q.operators",
    "qiskit_nature.second_q.hamiltonians", "qiskit_nature.second_q.problems",
    "qiskit_nature.second_q.operators.tensor_ordering",
]


def available() -> bool:
    return os.path.isfile(os.path.join(_OO_DIR, "partial_unitary_projection_optimizer.py"))


def _load(path: str, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load_reference():
    """Returns (PartialUnitaryProjectionOptimizer, BaseOptOrbSolver) classes of the reference."""
    if "v" in _cache:
        return _cache["v"]
    if not available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    pkg = types.ModuleType("_oo_ref_pkg")
    pkg.__path__ = [_OO_DIR]
    sys.modules["_oo_ref_pkg"] = pkg
    pupo = _load(os.path.join(_OO_DIR, "partial_unitary_projection_optimizer.py"),
                 "_oo_ref_pkg.partial_unitary_projection_optimizer")
    saved = {k: sys.modules.get(k) for k in _STUBS}
    try:
        for k in _STUBS:
            sys.modules[k] = mock.MagicMock()

        class _VariationalResult:  # real base class for BaseOptOrbResult
            def __init__(self) -> None:
                pass

        sys.modules["qiskit_algorithms.variational_algorithm"].VariationalResult = _VariationalResult
        base = _load(os.path.join(_OO_DIR, "base_opt_orb_solver.py"),
                     "_oo_ref_pkg.base_opt_orb_solver")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache["v"] = (pupo.PartialUnitaryProjectionOptimizer, base.BaseOptOrbSolver)
    return _cache["v"]


def make_solver(wavefunction_real: bool = True, weight_vector=None):
    """A BaseOptOrbSolver instance without running its qiskit-dependent constructor.  When
    ``weight_vector`` is given the object also carries a restated
    ``compute_rotated_weighted_energy_sum`` (opt_orb_eigensolver.py:149-169 cannot be imported:
    its module imports qiskit and the sibling eigensolvers)."""
    import torch

    _, BaseOptOrbSolver = load_reference()
    solver = BaseOptOrbSolver.__new__(BaseOptOrbSolver)
    solver.wavefunction_real = wavefunction_real
    if weight_vector is not None:
        solver.weight_vector = list(weight_vector)

        def compute_rotated_weighted_energy_sum(partial_unitary, oneRDM, twoRDM,
                                                one_body_integrals, two_body_integrals):
            total = 0
            for idx, (d, g) in enumerate(zip(oneRDM, twoRDM)):
                total += torch.tensor(solver.weight_vector[idx], dtype=torch.float64) * \
                    solver.compute_rotated_energy(partial_unitary=partial_unitary, oneRDM=d,
                                                  twoRDM=g, one_body_integrals=one_body_integrals,
                                                  two_body_integrals=two_body_integrals)
            return total

        solver.compute_rotated_weighted_energy_sum = compute_rotated_weighted_energy_sum
    return solver
