"""OptOrb outer loop with the B200 orbital optimiser (qiskit-free demo).

Mirrors examples/H2_OptOrbVQE.py / H4_OptOrbVQE.py of the reference: an eigensolver in a small
active space alternates with the orbital optimisation over M x N partial unitaries.  pyscf and
qiskit are not available in this environment, so the molecule is replaced by synthetic
molecule-like integrals and VQE by an exact diagonalisation (esoo_b200.harness); the orbital
optimiser is the drop-in CUDA class, used exactly as the reference class would be.

    python examples/optorb_fci_demo.py [--M 56] [--N 4] [--electrons 4]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import esoo_b200  # noqa: E402
from esoo_b200 import harness, synthetic  # noqa: E402


def molecule_like(M, seed=0):
    gen = torch.Generator().manual_seed(1000 + seed)
    eps = torch.linspace(-1.5, 1.0, M, dtype=torch.float64)
    noise = 0.1 * torch.randn(M, M, generator=gen, dtype=torch.float64)
    h = torch.diag(eps) + 0.5 * (noise + noise.T)
    g = synthetic.eri_spatial(M, seed=synthetic.SEED_ERI + seed, rank=12, scale=0.6)
    return synthetic.spin_orbital_integrals(h, g, "abba")     # the reference's (2M)^4 layout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=56)
    ap.add_argument("--N", type=int, default=4)
    ap.add_argument("--electrons", type=int, default=4)
    args = ap.parse_args()
    hs, gs = molecule_like(args.M)
    inner_iterations = []
    optimizer = esoo_b200.PartialUnitaryProjectionOptimizer(
        initial_BBstepsize=1e-3, stopping_tolerance=1e-9, maxiter=5000, device="cuda:0",
        callback=lambda it, e: inner_iterations.append(it))

    def outer_cb(it, energies, U):
        print(f"outer iteration {it}: E = {energies[0]:.10f}  (inner iterations so far: "
              f"{len(inner_iterations)})", flush=True)

    t0 = time.time()
    res = harness.run_outer_loop(optimizer, hs, gs, 2 * args.N, args.electrons // 2,
                                 args.electrons - args.electrons // 2, maxiter=10,
                                 stopping_tolerance=1e-8, outer_loop_callback=outer_cb)
    print(f"final energy {res['energies'][-1][0]:.10f} after {res['outer_iterations']} outer "
          f"iterations, {time.time() - t0:.2f} s")


if __name__ == "__main__":
    main()
