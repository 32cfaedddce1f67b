"""H2 (or a linear H_n chain) in 6-31G through the OptOrb outer loop with the B200 optimiser.

Counterpart of the reference's examples/H2_OptOrbVQE.py (H2, 0.735 Angstrom, 6-31G, reduced to 4
spin orbitals) and of its tests (tests/test_optorbvqe.py: expected energy -1.8661038 Ha).  The
6-31G hydrogen basis holds s functions only, so the MO integrals are built in closed form
(esoo_b200.molecule: Gaussian integrals + RHF) instead of with pyscf; VQE is replaced by an exact
diagonalisation in the active space (esoo_b200.harness).

    python examples/h2_631g_optorb.py [--atoms 2] [--N 2] [--states 1]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import esoo_b200  # noqa: E402
from esoo_b200 import harness, molecule, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--atoms", type=int, default=2)
    ap.add_argument("--spacing", type=float, default=0.735, help="Angstrom")
    ap.add_argument("--N", type=int, default=None, help="spatial orbitals kept (default: atoms)")
    ap.add_argument("--states", type=int, default=1)
    args = ap.parse_args()
    N = args.N or args.atoms
    mol = molecule.hydrogen_chain(args.atoms, args.spacing)
    M = mol["h"].shape[0]
    print(f"H{args.atoms} / 6-31G: M = {M} spatial orbitals -> N = {N}; "
          f"E_HF = {mol['e_hf']:.8f} (electronic), E_nuc = {mol['e_nuc']:.8f}")
    hs, gs = synthetic.spin_orbital_integrals(mol["h"], mol["g"], "abba")   # reference layout
    optimizer = esoo_b200.PartialUnitaryProjectionOptimizer(
        initial_BBstepsize=1e-3, stopping_tolerance=1e-5, maxiter=10000, device="cuda:0")

    def outer_cb(it, energies, U):
        print(f"outer iteration {it}: E = " + ", ".join(f"{e:.8f}" for e in energies), flush=True)

    t0 = time.time()
    res = harness.run_outer_loop(optimizer, hs, gs, 2 * N, mol["n_alpha"], mol["n_beta"],
                                 maxiter=20, stopping_tolerance=1e-5, n_states=args.states,
                                 outer_loop_callback=outer_cb,
                                 # rotated Hamiltonian (base_opt_orb_solver.py:597-604) from the
                                 # optimiser's device-resident engine instead of the CPU einsums
                                 engine_for_transform=optimizer)
    final = res["energies"][-1]
    print("final electronic energy " + ", ".join(f"{e:.8f}" for e in final) +
          f"  (total {final[0] + mol['e_nuc']:.8f}) after {res['outer_iterations']} outer "
          f"iterations, {time.time() - t0:.2f} s")
    if args.atoms == 2 and N == 2 and args.states == 1 and abs(args.spacing - 0.735) < 1e-12:
        print(f"reference test value -1.86610381: difference {final[0] + 1.8661038079694765:+.2e}")


if __name__ == "__main__":
    main()
