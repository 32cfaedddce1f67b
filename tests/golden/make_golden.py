"""Freeze golden vectors from the LIVE reference implementation (run in the build container only).

    python tests/golden/make_golden.py

For every case the inputs are seeded synthetic tensors (esoo_b200.synthetic) embedded in the
reference's spin-orbital layout; the outputs come from the unmodified reference code loaded from
/root/reference by oracle/ref_loader.py:
    E        BaseOptOrbSolver.compute_rotated_energy            base_opt_orb_solver.py:534-582
    grad     PartialUnitaryProjectionOptimizer.compute_rotated_energy_automatic_gradient  pupo.py:85-103
    orth     PartialUnitaryProjectionOptimizer.orth             pupo.py:70-83
    opt_*    PartialUnitaryProjectionOptimizer.compute_optimal_rotation  pupo.py:161-350
Small cases store their inputs as well, so the fixtures do not depend on the generator staying
bit-stable; large cases store input checksums.
"""
import os
import sys
from functools import partial

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import esoo_b200  # noqa: E402
from esoo_b200 import synthetic  # noqa: E402
from oracle import ref_loader  # noqa: E402

torch.set_num_threads(8)

CASES = [
    # name, M, N, pattern, n_states, weights, run_opt, (bb0, tol, maxiter), store_inputs, g_scale
    ("abba_M6_N2", 6, 2, "abba", 1, None, True, (0.05, 1e-9, 300), True, 1.0),
    ("abab_M5_N2", 5, 2, "abab", 1, None, True, (0.05, 1e-9, 300), True, 1.0),
    ("abba_M8_N3", 8, 3, "abba", 1, None, True, (0.02, 1e-8, 400), True, 1.0),
    ("weighted_M6_N2_k3", 6, 2, "abba", 3, [3, 2, 1], True, (0.02, 1e-9, 300), True, 1.0),
    ("abba_M12_N4", 12, 4, "abba", 1, None, True, (0.02, 1e-8, 500), True, 1.0),
    ("cfg1_M28_N2", 28, 2, "abba", 1, None, True, (0.02, 1e-8, 150), False, 1.0),
    ("cfg3_M20_N2_k3", 20, 2, "abba", 3, [3, 2, 1], True, (0.02, 1e-8, 150), False, 1.0),
    ("cfg2_M56_N4", 56, 4, "abba", 1, None, True, (0.02, 1e-8, 150), False, 1.0),
]

# BASELINE.json config 3 at its true shape (H2 cc-pV5Z: M=110 spatial -> 4 spin orbitals, 3
# state-averaged RDMs, weights [3,2,1]): an 18.7 GB spin-orbital tensor, run only on request
# (`make_golden.py big`, ~25 GB of host memory, a few minutes).
BIG_CASES = [
    ("cfg3_M110_N2_k3", 110, 2, "abba", 3, [3, 2, 1], True, (0.02, 1e-8, 40), False, 1.0),
]


def build_inputs(M, N, pattern, n_states, seed_shift=0, g_scale=1.0):
    h = synthetic.h_spatial(M, synthetic.SEED_H + seed_shift)
    g = synthetic.eri_spatial(M, synthetic.SEED_ERI + seed_shift, scale=g_scale)
    hs, gs = synthetic.spin_orbital_integrals(h, g, pattern)
    Ds, Gs = [], []
    for n in range(n_states):
        D, G = synthetic.rdms_spin(N, synthetic.SEED_RDM + seed_shift + 17 * n)
        Ds.append(D)
        Gs.append(G)
    U0 = synthetic.random_partial_unitary(M, N, synthetic.SEED_U + seed_shift)
    return hs, gs, Ds, Gs, U0


def main(cases=None, only=None):
    Pupo, _ = ref_loader.load_reference()
    for (name, M, N, pattern, k, weights, run_opt, optp, store, g_scale) in (cases or CASES):
        if only and name not in only:
            continue
        hs, gs, Ds, Gs, U0 = build_inputs(M, N, pattern, k, g_scale=g_scale)
        solver = ref_loader.make_solver(True, weights)
        if weights is None:
            fun, d_arg, g_arg = solver.compute_rotated_energy, Ds[0], Gs[0]
        else:
            fun, d_arg, g_arg = solver.compute_rotated_weighted_energy_sum, Ds, Gs
        obj = partial(fun, oneRDM=d_arg, twoRDM=g_arg, one_body_integrals=hs,
                      two_body_integrals=gs)
        opt = Pupo(initial_BBstepsize=0.1, stopping_tolerance=1e-6, maxiter=10)
        E = float(obj(partial_unitary=U0.clone()))
        grad = opt.compute_rotated_energy_automatic_gradient(U0.clone(), obj).numpy()
        V = U0 + 0.3 * synthetic.random_partial_unitary(M, N, 99)
        orthV = opt.orth(V).numpy()
        out = {"M": M, "N": N, "pattern": pattern, "n_states": k,
               "weights": np.array(weights if weights else [1.0], dtype=np.float64),
               "E": E, "grad": grad, "V": V.numpy(), "orthV": orthV, "U0": U0.numpy(),
               "checksum_g": float(gs.sum()), "checksum_h": float(hs.sum()),
               "checksum_G": float(sum(float(G.sum()) for G in Gs))}
        if store:
            out["h_spin"] = hs.numpy()
            out["g_spin"] = gs.numpy()
            for n in range(k):
                out[f"D_spin_{n}"] = Ds[n].numpy()
                out[f"G_spin_{n}"] = Gs[n].numpy()
        if run_opt:
            bb0, tol, maxiter = optp
            calls = []
            o2 = Pupo(initial_BBstepsize=bb0, stopping_tolerance=tol, maxiter=maxiter,
                      callback=lambda it, e: calls.append((it, e)))
            U_fin, E_fin = o2.compute_optimal_rotation(
                fun=fun, initial_partial_unitary=U0.clone(), oneRDM=d_arg, twoRDM=g_arg,
                one_body_integrals=hs, two_body_integrals=gs)
            out.update({"opt_bb0": bb0, "opt_tol": tol, "opt_maxiter": maxiter,
                        "opt_U": U_fin.numpy(), "opt_E": float(E_fin),
                        "opt_calls_it": np.array([c[0] for c in calls], dtype=np.int64),
                        "opt_calls_E": np.array([c[1] for c in calls], dtype=np.float64),
                        "opt_stepsize": float(o2.BBstepsize)})
            print(f"{name}: E={E:.12f} opt_E={float(E_fin):.12f} callbacks={len(calls)} "
                  f"last_it={calls[-1][0]}")
        else:
            print(f"{name}: E={E:.12f}")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


# ------------------------------------------------------------------------------------------------
# outer-loop goldens: the exact-diagonalisation harness (esoo_b200.harness) driven by the LIVE
# reference optimiser; the GPU test drives the same harness with the CUDA optimiser.
# ------------------------------------------------------------------------------------------------
OUTER_CASES = [
    # name, M, N, n_alpha, n_beta, n_states, weights, outer maxiter, outer tol, (bb0, tol, maxiter)
    ("outer_H2like_M8_N2", 8, 2, 1, 1, 1, None, 8, 1e-9, (1e-3, 1e-9, 2000)),
    ("outer_H4like_M12_N4", 12, 4, 2, 2, 1, None, 6, 1e-9, (1e-3, 1e-9, 3000)),
    ("outer_M10_N3", 10, 3, 2, 1, 1, None, 6, 1e-9, (1e-3, 1e-9, 3000)),
    ("outer_excited_M8_N2_k2", 8, 2, 1, 1, 2, [2, 1], 6, 1e-9, (1e-3, 1e-9, 2000)),
]


def molecule_like(M, seed=0):
    """Orbital-energy-like diagonal + weak couplings, PSD 8-fold-symmetric ERIs."""
    gen = torch.Generator().manual_seed(1000 + seed)
    eps = torch.linspace(-1.5, 1.0, M, dtype=torch.float64)
    noise = 0.1 * torch.randn(M, M, generator=gen, dtype=torch.float64)
    h = torch.diag(eps) + 0.5 * (noise + noise.T)
    g = synthetic.eri_spatial(M, seed=synthetic.SEED_ERI + seed, rank=12, scale=0.6)
    return synthetic.spin_orbital_integrals(h, g, "abba")


def main_outer():
    from esoo_b200 import harness
    Pupo, _ = ref_loader.load_reference()
    for (name, M, N, na, nb, k, weights, omax, otol, (bb0, tol, imax)) in OUTER_CASES:
        hs, gs = molecule_like(M, seed=M + N)
        solver = ref_loader.make_solver(True, weights)
        opt = Pupo(initial_BBstepsize=bb0, stopping_tolerance=tol, maxiter=imax)
        res = harness.run_outer_loop(opt, hs, gs, 2 * N, na, nb, maxiter=omax,
                                     stopping_tolerance=otol, n_states=k, weights=weights,
                                     energy_impl=solver.compute_rotated_energy)
        E = np.array(res["energies"], dtype=np.float64)
        print(f"{name}: outer iterations {len(E)}, energies {E[:, 0]}")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), M=M, N=N, n_alpha=na, n_beta=nb,
                            n_states=k, weights=np.array(weights if weights else [1.0]),
                            outer_maxiter=omax, outer_tol=otol, bb0=bb0, tol=tol, maxiter=imax,
                            h_spin=hs.numpy(), g_spin=gs.numpy(), energies=E,
                            U_final=res["U"].numpy())


# ------------------------------------------------------------------------------------------------
# the reference's OWN test system: H2, 0.735 Angstrom, 6-31G, 4 spin orbitals (M = 4, N = 2), with
# the optimiser settings of its tests.  `ref_test_golden` is the number hard-coded in the
# reference's test file; the run below (exact diagonalisation instead of VQE, live reference
# optimiser) lands on it, which pins integrals, conventions, harness and optimiser end to end.
# ------------------------------------------------------------------------------------------------
MOLECULE_CASES = [
    # name, atoms, N, n_states, weights, reference test golden (file:line)
    ("outer_H2_631G_ground", 2, 2, 1, None, [-1.8661038079694765]),     # tests/test_optorbvqe.py:67
    ("outer_H2_631G_k2", 2, 2, 2, [2, 1], [-1.85703467, -1.46615986]),  # tests/test_optorbmcvqe.py:61
    ("outer_H4_631G_ground", 4, 4, 1, None, None),                      # no golden: H4 chain, M = 8
]


def main_molecule():
    from esoo_b200 import harness, molecule
    Pupo, _ = ref_loader.load_reference()
    for (name, atoms, N, k, weights, ref_gold) in MOLECULE_CASES:
        mol = molecule.hydrogen_chain(atoms, 0.735)
        hs, gs = synthetic.spin_orbital_integrals(mol["h"], mol["g"], "abba")
        M = mol["h"].shape[0]
        solver = ref_loader.make_solver(True, weights)
        bb0, tol, imax, omax, otol = 1e-3, 1e-5, 10000, 20, 1e-5     # tests/test_optorbvqe.py:72-90
        opt = Pupo(initial_BBstepsize=bb0, stopping_tolerance=tol, maxiter=imax)
        res = harness.run_outer_loop(opt, hs, gs, 2 * N, atoms // 2, atoms // 2, maxiter=omax,
                                     stopping_tolerance=otol, n_states=k, weights=weights,
                                     energy_impl=solver.compute_rotated_energy)
        E = np.array(res["energies"], dtype=np.float64)
        print(f"{name}: outer iterations {len(E)}, final {E[-1]}, reference test golden {ref_gold}")
        extra = {} if ref_gold is None else {"ref_test_golden": np.array(ref_gold)}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), M=M, N=N, n_alpha=atoms // 2,
                            n_beta=atoms // 2, n_states=k,
                            weights=np.array(weights if weights else [1.0]),
                            outer_maxiter=omax, outer_tol=otol, bb0=bb0, tol=tol, maxiter=imax,
                            h_spin=hs.numpy(), g_spin=gs.numpy(), energies=E,
                            U_final=res["U"].numpy(), e_nuc=mol["e_nuc"], e_hf=mol["e_hf"], **extra)


# ------------------------------------------------------------------------------------------------
# compute_updated_partial_unitary (pupo.py:129-159) called directly on the live reference:
# iteration 0 keeps the step, odd / even iterations use the two Barzilai-Borwein formulas.
# ------------------------------------------------------------------------------------------------
def main_bb():
    Pupo, _ = ref_loader.load_reference()
    M, N, bb0 = 14, 3, 0.07
    gen = torch.Generator().manual_seed(4242)
    Uc = torch.linalg.qr(torch.randn(M, N, generator=gen, dtype=torch.float64))[0]
    Up = torch.linalg.qr(torch.randn(M, N, generator=gen, dtype=torch.float64))[0]
    Gc = torch.randn(M, N, generator=gen, dtype=torch.float64)
    Gp = torch.randn(M, N, generator=gen, dtype=torch.float64)
    out = {"M": M, "N": N, "bb0": bb0, "U_cur": Uc.numpy(), "U_prev": Up.numpy(),
           "G_cur": Gc.numpy(), "G_prev": Gp.numpy(), "iterations": np.array([0, 1, 2, 5, 6])}
    for it in out["iterations"]:
        opt = Pupo(initial_BBstepsize=bb0, stopping_tolerance=1e-6, maxiter=10)
        U_next = opt.compute_updated_partial_unitary(int(it), Uc.clone(), Up.clone(), Gc.clone(),
                                                     Gp.clone())
        out[f"U_next_{it}"] = U_next.numpy()
        out[f"step_{it}"] = float(opt.BBstepsize)
        print(f"bb_update iteration {it}: step {float(opt.BBstepsize):.12e}")
    np.savez_compressed(os.path.join(HERE, "bb_update_M14_N3.npz"), **out)


# ------------------------------------------------------------------------------------------------
# a non-default decay_factor changes the smoothed stopping measure (pupo.py:230,268,320) and hence
# the iteration at which the loop stops
# ------------------------------------------------------------------------------------------------
def main_decay():
    Pupo, _ = ref_loader.load_reference()
    M, N = 6, 2
    hs, gs, Ds, Gs, U0 = build_inputs(M, N, "abba", 1)
    solver = ref_loader.make_solver(True, None)
    out = {"M": M, "N": N, "pattern": "abba", "n_states": 1, "weights": np.array([1.0]),
           "h_spin": hs.numpy(),
           "g_spin": gs.numpy(), "D_spin_0": Ds[0].numpy(), "G_spin_0": Gs[0].numpy(),
           "U0": U0.numpy(), "bb0": 0.05, "tol": 1e-7, "maxiter": 300,
           "decays": np.array([0.2, 0.5, 0.95])}
    for d in out["decays"]:
        calls = []
        opt = Pupo(initial_BBstepsize=0.05, stopping_tolerance=1e-7, maxiter=300,
                   decay_factor=float(d), callback=lambda it, e: calls.append((it, e)))
        U_fin, E_fin = opt.compute_optimal_rotation(
            fun=solver.compute_rotated_energy, initial_partial_unitary=U0.clone(), oneRDM=Ds[0],
            twoRDM=Gs[0], one_body_integrals=hs, two_body_integrals=gs)
        out[f"E_{d}"] = float(E_fin)
        out[f"U_{d}"] = U_fin.numpy()
        out[f"calls_it_{d}"] = np.array([c[0] for c in calls], dtype=np.int64)
        print(f"decay {d}: {len(calls)} callbacks, E = {float(E_fin):.12f}")
    np.savez_compressed(os.path.join(HERE, "opt_decay_M6_N2.npz"), **out)


# ------------------------------------------------------------------------------------------------
# gradient_method='finite_difference' through compute_optimal_rotation (pupo.py:105-127,183-184):
# the live reference's trajectory with its own central-difference gradient (step 1e-8).
# ------------------------------------------------------------------------------------------------
def main_fd():
    Pupo, _ = ref_loader.load_reference()
    M, N = 5, 2
    hs, gs, Ds, Gs, U0 = build_inputs(M, N, "abab", 1)
    solver = ref_loader.make_solver(True, None)
    calls = []
    opt = Pupo(initial_BBstepsize=0.05, stopping_tolerance=1e-7, maxiter=12,
               gradient_method='finite_difference', callback=lambda it, e: calls.append((it, e)))
    U_fin, E_fin = opt.compute_optimal_rotation(
        fun=solver.compute_rotated_energy, initial_partial_unitary=U0.clone(), oneRDM=Ds[0],
        twoRDM=Gs[0], one_body_integrals=hs, two_body_integrals=gs)
    obj = partial(solver.compute_rotated_energy, oneRDM=Ds[0], twoRDM=Gs[0],
                  one_body_integrals=hs, two_body_integrals=gs)
    fd_grad = Pupo(0.05, 1e-7, 12).compute_rotated_energy_gradient(U0.clone(), obj)
    print(f"fd_M5_N2: {len(calls)} callbacks, E = {float(E_fin):.12f}")
    np.savez_compressed(
        os.path.join(HERE, "fd_M5_N2.npz"), M=M, N=N, pattern="abab", n_states=1,
        weights=np.array([1.0]), h_spin=hs.numpy(), g_spin=gs.numpy(), D_spin_0=Ds[0].numpy(),
        G_spin_0=Gs[0].numpy(), U0=U0.numpy(), bb0=0.05, tol=1e-7, maxiter=12,
        opt_U=U_fin.numpy(), opt_E=float(E_fin), opt_stepsize=float(opt.BBstepsize),
        opt_calls_it=np.array([c[0] for c in calls], dtype=np.int64),
        opt_calls_E=np.array([float(c[1]) for c in calls], dtype=np.float64),
        fd_grad_U0=fd_grad.numpy())


# ------------------------------------------------------------------------------------------------
# tensor part of get_rotated_hamiltonian (base_opt_orb_solver.py:597-604): the two einsums exactly
# as the reference writes them, for both spin patterns and an odd M.
# ------------------------------------------------------------------------------------------------
def main_rotated():
    out = {}
    for tag, M, N, pattern in (("a", 6, 2, "abba"), ("b", 7, 3, "abab"), ("c", 12, 4, "abba")):
        hs, gs, _, _, U0 = build_inputs(M, N, pattern, 1, seed_shift=3)
        W = torch.block_diag(U0, U0)
        h_rot = torch.einsum('pq,pi,qj->ij', hs, W, W)
        g_rot = torch.einsum('pqrs,pi,qj,rk,sl->ijkl', gs, W, W, W, W)
        out.update({f"{tag}_M": M, f"{tag}_N": N, f"{tag}_pattern": pattern,
                    f"{tag}_h_spin": hs.numpy(), f"{tag}_g_spin": gs.numpy(), f"{tag}_U": U0.numpy(),
                    f"{tag}_h_rot": h_rot.numpy(), f"{tag}_g_rot": g_rot.numpy()})
        print(f"rotated {tag}: M={M} N={N} {pattern} |g'|={float(g_rot.abs().sum()):.6f}")
    np.savez_compressed(os.path.join(HERE, "rotated_integrals.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) < 2 or sys.argv[1] == "fd":
        main_fd()
    if len(sys.argv) < 2 or sys.argv[1] == "rotated":
        main_rotated()
    if len(sys.argv) < 2 or sys.argv[1] == "decay":
        main_decay()
    if len(sys.argv) < 2 or sys.argv[1] == "bb":
        main_bb()
    if len(sys.argv) < 2 or sys.argv[1] == "molecule":
        main_molecule()
    if len(sys.argv) < 2 or sys.argv[1] == "inner":
        main(only=sys.argv[2:])
    if len(sys.argv) >= 2 and sys.argv[1] == "big":
        main(BIG_CASES)
    if len(sys.argv) < 2 or sys.argv[1] == "outer":
        main_outer()
