"""The numpy oracle against golden vectors frozen from the live reference (CPU only)."""
import numpy as np
import pytest

from conftest import golden_inputs, golden_inputs_spatial, golden_names, load_golden
from oracle import oracle_np as onp

BIG_M = 60        # above this the (2M)^4 embedding is skipped on the CPU (cfg 3: 18.7 GB)


@pytest.mark.parametrize("name", golden_names())
def test_energy_and_gradient(name):
    gold = load_golden(name)
    if int(gold["M"]) > 30:
        pytest.skip("spin-orbital oracle at M=56 is slow; covered in the spatial test")
    hs, gs, Ds, Gs, U0 = golden_inputs(gold)
    w = gold["weights"]
    U = U0.numpy()
    args = ([d.numpy() for d in Ds], [g.numpy() for g in Gs], hs.numpy(), gs.numpy(), w)
    E = onp.weighted_energy_sum_spin(U, *args)
    grad = onp.weighted_energy_grad_spin(U, *args)
    assert abs(E - float(gold["E"])) <= 1e-11 * max(1.0, abs(float(gold["E"])))
    rel = np.linalg.norm(grad - gold["grad"]) / np.linalg.norm(gold["grad"])
    assert rel <= 1e-11


@pytest.mark.parametrize("name", golden_names())
def test_spatial_reduction_matches_reference(name):
    """The spin->spatial reduction of the product's ingest + the spatial oracle reproduce the
    reference's spin-orbital energy and autograd gradient."""
    import torch
    from esoo_b200 import ingest

    gold = load_golden(name)
    if int(gold["M"]) > BIG_M:
        h, g, D, G, U0 = golden_inputs_spatial(gold)
    else:
        hs, gs, Ds, Gs, U0 = golden_inputs(gold)
        h, g, st = ingest.reduce_integrals(hs, gs)
        D, G = ingest.reduce_rdms(Ds, Gs, st, list(gold["weights"]))
        assert sorted(st.blocks) == sorted(
            [(s, t, t, s) if str(gold["pattern"]) == "abba" else (s, t, s, t)
             for s in (0, 1) for t in (0, 1)])
    U = U0.numpy()
    E = onp.rotated_energy_spatial(U, D.numpy(), G.numpy(), h.numpy(), g.numpy())
    grad = onp.rotated_energy_grad_spatial(U, D.numpy(), G.numpy(), h.numpy(), g.numpy())
    assert abs(E - float(gold["E"])) <= 1e-11 * max(1.0, abs(float(gold["E"])))
    rel = np.linalg.norm(grad - gold["grad"]) / np.linalg.norm(gold["grad"])
    assert rel <= 1e-11


@pytest.mark.parametrize("name", golden_names())
def test_orth(name):
    gold = load_golden(name)
    out = onp.orth(gold["V"])
    assert np.max(np.abs(out - gold["orthV"])) <= 1e-12


@pytest.mark.parametrize("name", [n for n in golden_names() if "opt_E" in load_golden(n)])
def test_optimal_rotation_trajectory(name):
    """Driver loop, BB step and stopping rule: same callbacks, iteration count and final U.
    Measured deviations from the live reference over all fixtures (up to 445 BB steps): callback
    energies <= 1.9e-12 relative, final U <= 4.9e-11, final energy <= 2.2e-14; the bounds below
    are 10x those."""
    gold = load_golden(name)
    from esoo_b200 import ingest
    big = int(gold["M"]) > BIG_M
    if big:
        h, g, D, G, U0 = golden_inputs_spatial(gold)
    else:
        hs, gs, Ds, Gs, U0 = golden_inputs(gold)
        h, g, st = ingest.reduce_integrals(hs, gs)
        D, G = ingest.reduce_rdms(Ds, Gs, st, list(gold["weights"]))
    hn, gn, Dn, Gn = h.numpy(), g.numpy(), D.numpy(), G.numpy()
    # cfg 3 at its true size: one oracle gradient takes seconds, so only the three hand-unrolled
    # iterations are replayed (the trajectory prefix does not depend on maxiter)
    maxiter = 2 if big else int(gold["opt_maxiter"])
    res = onp.optimal_rotation(
        lambda U: onp.rotated_energy_spatial(U, Dn, Gn, hn, gn),
        lambda U: onp.rotated_energy_grad_spatial(U, Dn, Gn, hn, gn),
        U0.numpy(), float(gold["opt_bb0"]), float(gold["opt_tol"]), maxiter)
    its = [c[0] for c in res["callbacks"]]
    Es = np.array([c[1] for c in res["callbacks"]])
    scale = max(1.0, float(np.max(np.abs(gold["opt_calls_E"]))))
    if big:
        assert its == list(gold["opt_calls_it"][:len(its)]) and len(its) == 3
        assert np.max(np.abs(Es - gold["opt_calls_E"][:3])) <= 1e-10 * scale
        return
    assert its == list(gold["opt_calls_it"])
    assert abs(res["energy"] - float(gold["opt_E"])) <= 1e-12 * scale
    assert np.max(np.abs(Es - gold["opt_calls_E"])) <= 2e-11 * scale
    assert np.max(np.abs(res["U"] - gold["opt_U"])) <= 5e-10


# ---------------------------------------------------------------------------------------------
# The reference's own test system (H2, 0.735 Angstrom, 6-31G -> 4 spin orbitals) and the numbers
# hard-coded in its test files.  pyscf / qiskit are absent, so the integrals come from the closed
# forms in esoo_b200.molecule and the eigensolver from the exact-diagonalisation harness.
# ---------------------------------------------------------------------------------------------
def test_h2_631g_integrals():
    from esoo_b200 import molecule, synthetic, harness
    mol = molecule.hydrogen_chain(2, 0.735)
    h, g = mol["h"].numpy(), mol["g"].numpy()
    assert h.shape == (4, 4) and g.shape == (4, 4, 4, 4)
    assert abs(mol["e_nuc"] - 0.52917721092 / 0.735) < 1e-12
    # literature values for H2 / 6-31G at 0.735 Angstrom: RHF -1.1268, FCI -1.1516 (total energies)
    assert abs(mol["e_hf"] + mol["e_nuc"] + 1.12681) < 5e-5
    # MO basis: Fock matrix diagonal => Brillouin: h is not diagonal, but the 8-fold symmetry of
    # (pq|rs) must hold for g[p,q,r,s] = -1/2 (ps|qr)
    eri = -2.0 * g.transpose(0, 3, 1, 2)                     # (ps|qr) -> eri[p,s,q,r]
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (2, 3, 0, 1)]:
        assert np.max(np.abs(eri - eri.transpose(perm))) < 1e-13
    # HF energy from the MO integrals: 2 h_00 + (00|00)
    assert abs(2 * h[0, 0] + eri[0, 0, 0, 0] - mol["e_hf"]) < 1e-10
    # full CI in all 8 spin orbitals
    hs, gs = synthetic.spin_orbital_integrals(mol["h"], mol["g"], "abba")
    sec = harness.FockSector(8, 2)
    sub = sec.sz_subspace(1, 4)
    H = sec.hamiltonian(hs.numpy(), gs.numpy())
    e_fci = np.linalg.eigvalsh(H[np.ix_(sub, sub)])[0]
    assert abs(e_fci + mol["e_nuc"] + 1.15161) < 5e-5
    assert e_fci < mol["e_hf"]


REF_TEST_CASES = ["outer_H2_631G_ground", "outer_H2_631G_k2"]


@pytest.mark.parametrize("name", REF_TEST_CASES)
def test_fixture_matches_reference_test_golden(name):
    """The live-reference run frozen in the fixture reproduces the number hard-coded in the
    reference's own tests (tests/test_optorbvqe.py:67,98-100 and tests/test_optorbmcvqe.py:61,96-98;
    their assertion is decimal=3, i.e. 1.5e-3)."""
    gold = load_golden(name)
    final, ref = gold["energies"][-1], gold["ref_test_golden"]
    assert np.max(np.abs(final - ref)) < 1.5e-3            # the reference's own bar
    assert np.max(np.abs(final - ref)) < 2e-5              # what is actually achieved
    # the stored integrals are what esoo_b200.molecule produces today
    from esoo_b200 import molecule, synthetic
    mol = molecule.hydrogen_chain(2, 0.735)
    hs, gs = synthetic.spin_orbital_integrals(mol["h"], mol["g"], "abba")
    assert np.max(np.abs(hs.numpy() - gold["h_spin"])) < 1e-12
    assert np.max(np.abs(gs.numpy() - gold["g_spin"])) < 1e-12


class _OracleOptimizer:
    """The optimiser protocol of the reference (device, compute_optimal_rotation keyword call)
    served by the numpy oracle; lets the harness run without the reference tree."""
    device = "cpu"

    def __init__(self, bb0, tol, maxiter):
        self.bb0, self.tol, self.maxiter = bb0, tol, maxiter

    def compute_optimal_rotation(self, fun, initial_partial_unitary, oneRDM, twoRDM,
                                 one_body_integrals, two_body_integrals):
        import torch
        hs, gs = one_body_integrals.numpy(), two_body_integrals.numpy()
        if isinstance(oneRDM, list):
            w = list(fun.__self__.weight_vector)
            Ds, Gs = [d.numpy() for d in oneRDM], [g.numpy() for g in twoRDM]
            e = lambda U: onp.weighted_energy_sum_spin(U, Ds, Gs, hs, gs, w)
            gr = lambda U: onp.weighted_energy_grad_spin(U, Ds, Gs, hs, gs, w)
        else:
            D, G = oneRDM.numpy(), twoRDM.numpy()
            e = lambda U: onp.rotated_energy_spin(U, D, G, hs, gs)
            gr = lambda U: onp.rotated_energy_grad_spin(U, D, G, hs, gs)
        res = onp.optimal_rotation(e, gr, initial_partial_unitary.numpy(), self.bb0, self.tol,
                                   self.maxiter)
        return torch.from_numpy(res["U"]), torch.tensor(res["energy"], dtype=torch.float64)


@pytest.mark.parametrize("name", REF_TEST_CASES + ["outer_H4_631G_ground"])
def test_oracle_outer_loop_on_molecule(name):
    """Oracle optimiser inside the outer loop: every outer energy within 1e-8 Ha of the run with
    the live reference optimiser, hence on the reference's test goldens."""
    import torch
    from esoo_b200 import harness
    gold = load_golden(name)
    k, N = int(gold["n_states"]), int(gold["N"])
    weights = list(gold["weights"]) if k > 1 else None
    opt = _OracleOptimizer(float(gold["bb0"]), float(gold["tol"]), int(gold["maxiter"]))
    res = harness.run_outer_loop(opt, torch.from_numpy(gold["h_spin"]), torch.from_numpy(gold["g_spin"]),
                                 2 * N, int(gold["n_alpha"]), int(gold["n_beta"]),
                                 maxiter=int(gold["outer_maxiter"]),
                                 stopping_tolerance=float(gold["outer_tol"]), n_states=k,
                                 weights=weights)
    E = np.array(res["energies"])
    assert E.shape == gold["energies"].shape
    assert np.max(np.abs(E - gold["energies"])) <= 1e-8
    if "ref_test_golden" in gold:
        assert np.max(np.abs(E[-1] - gold["ref_test_golden"])) < 2e-5


def test_finite_difference_trajectory_vs_reference_golden():
    """gradient_method='finite_difference' (pupo.py:105-127, 183-184): the oracle's driver with a
    central-difference gradient (step 1e-8) follows the live reference's FD run.  FD gradients
    carry ~1e-8 relative noise that depends on summation order, so the trajectories agree to
    ~1e-6, not to machine precision."""
    gold = load_golden("fd_M5_N2")
    D, G, h, g = gold["D_spin_0"], gold["G_spin_0"], gold["h_spin"], gold["g_spin"]
    energy = lambda U: onp.rotated_energy_spin(U, D, G, h, g)

    def fd(U):
        out = np.empty_like(U)
        for i in range(U.shape[0]):
            for j in range(U.shape[1]):
                up, um = U.copy(), U.copy()
                up[i, j] += 1e-8
                um[i, j] -= 1e-8
                out[i, j] = (energy(up) - energy(um)) / 2e-8
        return out

    assert np.max(np.abs(fd(gold["U0"]) - gold["fd_grad_U0"])) <= 1e-6
    res = onp.optimal_rotation(energy, fd, gold["U0"], float(gold["bb0"]), float(gold["tol"]),
                               int(gold["maxiter"]))
    assert [c[0] for c in res["callbacks"]] == list(gold["opt_calls_it"])
    assert np.max(np.abs(np.array([c[1] for c in res["callbacks"]]) - gold["opt_calls_E"])) <= 1e-5
    assert abs(res["energy"] - float(gold["opt_E"])) <= 1e-5


def test_rotated_integrals_vs_reference_golden():
    """h', g' of get_rotated_hamiltonian (base_opt_orb_solver.py:597-604), frozen from the
    reference's own einsums: the oracle in both pictures, and the spin-block embedding
    (esoo_b200.rotated.expand_spin_blocks) the product uses after the CUDA transform."""
    import torch
    from esoo_b200 import ingest, rotated
    gold = load_golden("rotated_integrals")
    for tag in "abc":
        hs, gs, U = gold[f"{tag}_h_spin"], gold[f"{tag}_g_spin"], gold[f"{tag}_U"]
        h_rot, g_rot = onp.rotated_integrals_spin(U, hs, gs)
        assert np.max(np.abs(h_rot - gold[f"{tag}_h_rot"])) <= 1e-12
        assert np.max(np.abs(g_rot - gold[f"{tag}_g_rot"])) <= 1e-12
        h, g, st = ingest.reduce_integrals(torch.from_numpy(hs), torch.from_numpy(gs))
        h_sp, g_sp = onp.rotated_integrals_spatial(U, h.numpy(), g.numpy())
        he, ge = rotated.expand_spin_blocks(torch.from_numpy(h_sp), torch.from_numpy(g_sp), st)
        assert np.max(np.abs(he - gold[f"{tag}_h_rot"])) <= 1e-12
        assert np.max(np.abs(ge - gold[f"{tag}_g_rot"])) <= 1e-12


def test_torch_port_matches_reference_golden():
    """oracle/torch_port.py (the CPU baseline of bench.py) states the reference's einsum + autograd
    formulation: same E and dE/dU as the live reference, in the spin-orbital and in the spatial
    picture, and additive over slabs of the last ERI index (what the bounded sample relies on)."""
    import torch
    from esoo_b200 import ingest
    from oracle import torch_port
    for name in ("abba_M6_N2", "weighted_M6_N2_k3", "abab_M5_N2"):
        gold = load_golden(name)
        hs, gs, Ds, Gs, U0 = golden_inputs(gold)
        w = list(gold["weights"])
        _, _, E, grad = torch_port.time_reference_spin(U0, Ds, Gs, w, hs, gs)
        assert abs(E - float(gold["E"])) <= 1e-12 * max(1.0, abs(float(gold["E"])))
        assert np.linalg.norm(grad.numpy() - gold["grad"]) <= 1e-12 * np.linalg.norm(gold["grad"])
        h, g, st = ingest.reduce_integrals(hs, gs)
        D, G = ingest.reduce_rdms(Ds, Gs, st, w)
        E2, grad2 = torch_port.energy_and_autograd(U0, D, G, h, g, 0)
        assert abs(E2 - float(gold["E"])) <= 1e-12 * max(1.0, abs(float(gold["E"])))
        assert np.linalg.norm(grad2.numpy() - gold["grad"]) <= 1e-12 * np.linalg.norm(gold["grad"])
        M = h.shape[0]
        half = M // 2
        Ea, ga = torch_port.energy_and_autograd(U0, D, G, h, g[..., :half].contiguous(), 0)
        Eb, gb = torch_port.energy_and_autograd(U0, D, G, h, g[..., half:].contiguous(), half)
        assert abs(Ea + Eb - E2) <= 1e-12 * max(1.0, abs(E2))
        assert np.linalg.norm((ga + gb - grad2).numpy()) <= 1e-12 * np.linalg.norm(gold["grad"])


def test_bb_update_vs_reference_golden():
    """compute_updated_partial_unitary (pupo.py:129-159) of the live reference, iteration 0 / odd /
    even: next iterate and the Barzilai-Borwein step it leaves in BBstepsize."""
    gold = load_golden("bb_update_M14_N3")
    for it in gold["iterations"]:
        U_next, step = onp.bb_update(int(it), gold["U_cur"], gold["U_prev"], gold["G_cur"],
                                     gold["G_prev"], float(gold["bb0"]))
        assert abs(step - float(gold[f"step_{it}"])) <= 1e-13 * abs(step)
        assert np.max(np.abs(U_next - gold[f"U_next_{it}"])) <= 1e-12


def test_decay_factor_vs_reference_golden():
    """Non-default decay_factor: the smoothed stopping measure (pupo.py:230,268,320) stops the
    loop at a different iteration; the oracle follows the live reference for each value."""
    gold = load_golden("opt_decay_M6_N2")
    hs, gs, Ds, Gs, U0 = golden_inputs(gold)
    D, G, h, g = Ds[0].numpy(), Gs[0].numpy(), hs.numpy(), gs.numpy()
    counts = []
    for d in gold["decays"]:
        res = onp.optimal_rotation(lambda U: onp.rotated_energy_spin(U, D, G, h, g),
                                   lambda U: onp.rotated_energy_grad_spin(U, D, G, h, g),
                                   U0.numpy(), float(gold["bb0"]), float(gold["tol"]),
                                   int(gold["maxiter"]), decay_factor=float(d))
        assert [c[0] for c in res["callbacks"]] == list(gold[f"calls_it_{d}"])
        assert abs(res["energy"] - float(gold[f"E_{d}"])) <= 1e-8
        counts.append(len(res["callbacks"]))
    assert len(set(counts)) == len(counts)         # the parameter really changes the stopping point


def test_patch_reference_routes_rotated_hamiltonian(monkeypatch):
    """esoo_b200.patch_reference on the LIVE reference class (qiskit mocked, as in ref_loader): the
    patched BaseOptOrbSolver.get_rotated_hamiltonian hands ElectronicEnergy.from_raw_integrals the
    same h1_a / h2_aa as the reference's own CPU einsums (base_opt_orb_solver.py:597-612), with the
    tensors coming from the optimiser's engine (here an oracle-backed stand-in for the CUDA one);
    solvers whose optimiser is not an esoo_b200 one keep the original method."""
    import sys
    import types
    from unittest import mock
    import torch
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present on this machine")
    import esoo_b200
    from esoo_b200 import ingest
    _, Base = ref_loader.load_reference()
    base_mod = sys.modules[Base.__module__]
    gold = load_golden("rotated_integrals")
    hs, gs = torch.from_numpy(gold["c_h_spin"]), torch.from_numpy(gold["c_g_spin"])
    U = torch.from_numpy(gold["c_U"])
    N = int(gold["c_N"])

    class OracleEngine:                      # what OrbitalEngine.transform returns, from the oracle
        def __init__(self, h, g):
            self.h, self.g = h, g

        def transform(self, Ux):
            h_rot, g_rot = onp.rotated_integrals_spatial(Ux.numpy(), self.h.numpy(), self.g.numpy())
            return torch.from_numpy(h_rot), torch.from_numpy(g_rot)

    class FakeOptimizer:
        def _engine_for(self, h, g):
            h_sp, g_sp, st = ingest.reduce_integrals(h, g)
            return OracleEngine(h_sp, g_sp), st

    def make_solver(optimizer):
        sv = Base.__new__(Base)
        sv.one_body_integrals, sv.two_body_integrals = hs, gs
        sv.num_spin_orbitals = 2 * N
        sv.mapper = mock.MagicMock()
        sv._partial_unitary_optimizer_list = [None, optimizer]
        return sv

    captured = []
    fake_ee = mock.MagicMock()
    fake_ee.from_raw_integrals.side_effect = lambda h1_a, h2_aa: captured.append(
        (np.array(h1_a), np.array(h2_aa))) or mock.MagicMock()
    monkeypatch.setattr(base_mod, "ElectronicEnergy", fake_ee)
    ham_mod = types.ModuleType("qiskit_nature.second_q.hamiltonians")
    ham_mod.ElectronicEnergy = fake_ee
    monkeypatch.setitem(sys.modules, "qiskit_nature.second_q.hamiltonians", ham_mod)

    original = Base.get_rotated_hamiltonian
    make_solver(object()).get_rotated_hamiltonian(U)            # the reference's own einsums
    try:
        esoo_b200.patch_reference(Base)
        make_solver(FakeOptimizer()).get_rotated_hamiltonian(U)  # routed through the engine
        make_solver(object()).get_rotated_hamiltonian(U)         # foreign optimiser: original path
    finally:
        Base.get_rotated_hamiltonian = original
    assert len(captured) == 3
    (h_ref, g_ref), (h_new, g_new), (h_old, g_old) = captured
    assert h_ref.shape == (N, N) and g_ref.shape == (N, N, N, N)
    assert np.max(np.abs(h_new - h_ref)) <= 1e-12 and np.max(np.abs(g_new - g_ref)) <= 1e-12
    assert np.array_equal(h_old, h_ref) and np.array_equal(g_old, g_ref)
