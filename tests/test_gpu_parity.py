"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI of
liboo_b200.so, against (a) golden vectors frozen from the live reference, (b) the numpy oracle on
seeded inputs, (c) size-independent properties at the BASELINE.json size (M=256, N=16).

Tolerances are the ones BASELINE.json's north_star states:
    |dE| <= 1e-10 Ha per evaluation, relative dE/dU error <= 1e-9, final energies within 1e-8 Ha.
"""
from functools import partial

import numpy as np
import pytest

from conftest import (golden_inputs, golden_names, load_golden, record_deviation,
                      shard_partial_oracle)

pytestmark = pytest.mark.gpu

E_TOL = 1e-10
G_RTOL = 1e-9
EFINAL_TOL = 1e-8
# whole trajectories (up to 445 BB steps) against the live reference / the oracle: bounds = 10-20x
# the largest deviation measured on the B200 over all fixtures and over the kernel versions of the
# round (profiles/r02_parity_deviations.jsonl: callback energies 7.9e-12 .. 1.0e-11 relative, final
# U 3.0e-12 .. 5.4e-11, final energy 2.8e-14); round 1 accepted 1e-7 / 1e-5
TRAJ_E_RTOL = 2e-10
TRAJ_U_TOL = 1e-9


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("CUDA device required for -m gpu tests (no CPU fallback exists)")
    return torch


def _rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


class _Solver:
    """Stand-in for the reference's BaseOptOrbSolver / OptOrbEigensolver: the optimiser only needs
    the bound method's name and `weight_vector`."""

    def __init__(self, weights=None):
        self.wavefunction_real = True
        if weights is not None:
            self.weight_vector = list(weights)

    def compute_rotated_energy(self, *a, **k):
        raise AssertionError("the CUDA optimiser must not call the Python objective")

    def compute_rotated_weighted_energy_sum(self, *a, **k):
        raise AssertionError("the CUDA optimiser must not call the Python objective")


def _fun_and_args(gold, Ds, Gs):
    if int(gold["n_states"]) == 1:
        return _Solver().compute_rotated_energy, Ds[0], Gs[0]
    return _Solver(gold["weights"]).compute_rotated_weighted_energy_sum, Ds, Gs


# --------------------------------------------------------------------------------------------
# (a) golden vectors from the live reference
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_energy_gradient_vs_reference_golden(torch_cuda, name):
    import esoo_b200
    gold = load_golden(name)
    hs, gs, Ds, Gs, U0 = golden_inputs(gold)
    fun, d_arg, g_arg = _fun_and_args(gold, Ds, Gs)
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 10, device="cuda:0")
    obj = partial(fun, oneRDM=d_arg, twoRDM=g_arg, one_body_integrals=hs, two_body_integrals=gs)
    grad = opt.compute_rotated_energy_automatic_gradient(U0.clone(), obj).cpu().numpy()
    assert _rel(grad, gold["grad"]) <= G_RTOL
    eng = opt._prepare(fun, d_arg, g_arg, hs, gs, U0.shape[1])
    E, _ = eng.energy_grad_host(U0.numpy())
    assert abs(E - float(gold["E"])) <= E_TOL
    esoo_b200.clear_engine_cache()


@pytest.mark.parametrize("name", golden_names())
def test_orth_vs_reference_golden(torch_cuda, name):
    import esoo_b200
    torch = torch_cuda
    gold = load_golden(name)
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 10, device="cuda:0")
    out = opt.orth(torch.from_numpy(gold["V"])).cpu().numpy()
    assert np.max(np.abs(out - gold["orthV"])) <= 1e-12


@pytest.mark.parametrize("name", [n for n in golden_names() if "opt_E" in load_golden(n)])
def test_optimal_rotation_vs_reference_golden(torch_cuda, name):
    """Whole inner loop on the device: callbacks, iteration count, final U and energy."""
    import esoo_b200
    gold = load_golden(name)
    hs, gs, Ds, Gs, U0 = golden_inputs(gold)
    fun, d_arg, g_arg = _fun_and_args(gold, Ds, Gs)
    calls = []
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(
        float(gold["opt_bb0"]), float(gold["opt_tol"]), int(gold["opt_maxiter"]),
        callback=lambda it, e: calls.append((it, e)), device="cuda:0")
    U, E = opt.compute_optimal_rotation(fun=fun, initial_partial_unitary=U0.clone(),
                                        oneRDM=d_arg, twoRDM=g_arg, one_body_integrals=hs,
                                        two_body_integrals=gs)
    assert U.device.type == "cpu" and U.dtype == torch_cuda.float64 and E.dim() == 0
    scale = max(1.0, float(np.max(np.abs(gold["opt_calls_E"]))))
    dcb = float(np.max(np.abs(np.array([c[1] for c in calls]) - gold["opt_calls_E"]))) / scale
    dU = float(np.max(np.abs(U.numpy() - gold["opt_U"])))
    record_deviation("optimal_rotation:" + name, dE_final=abs(float(E) - float(gold["opt_E"])),
                     dE_callbacks_rel=dcb, dU=dU, iterations=len(calls))
    assert abs(float(E) - float(gold["opt_E"])) <= EFINAL_TOL
    assert [c[0] for c in calls] == list(gold["opt_calls_it"])
    assert dcb <= TRAJ_E_RTOL
    assert dU <= TRAJ_U_TOL
    UtU = U.numpy().T @ U.numpy()
    assert np.max(np.abs(UtU - np.eye(U.shape[1]))) <= 1e-12
    esoo_b200.clear_engine_cache()


# --------------------------------------------------------------------------------------------
# (b) numpy oracle on seeded inputs: every kernel template, padding, multi-pass, sharding
# --------------------------------------------------------------------------------------------
def _spatial_case(torch, M, N, seed=0):
    from esoo_b200 import synthetic
    h = synthetic.h_spatial(M, synthetic.SEED_H + seed)
    g = synthetic.eri_spatial(M, synthetic.SEED_ERI + seed)
    D, G = synthetic.rdms_spatial(N, synthetic.SEED_RDM + seed)
    U = synthetic.random_partial_unitary(M, N, synthetic.SEED_U + seed)
    return h, g, D, G, U


@pytest.mark.parametrize("M,N", [(8, 1), (10, 2), (16, 8), (23, 5), (40, 9), (48, 16), (36, 17),
                                 (56, 24), (40, 25), (64, 32), (57, 32)])
def test_energy_gradient_vs_oracle(torch_cuda, M, N):
    import esoo_b200
    from oracle import oracle_np as onp
    torch = torch_cuda
    h, g, D, G, U = _spatial_case(torch, M, N, seed=M + N)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h, g)
    eng.set_rdms(D, G)
    E, grad = eng.energy_grad(U)
    E_ref = onp.rotated_energy_spatial(U.numpy(), D.numpy(), G.numpy(), h.numpy(), g.numpy())
    g_ref = onp.rotated_energy_grad_spatial(U.numpy(), D.numpy(), G.numpy(), h.numpy(), g.numpy())
    assert abs(float(E) - E_ref) <= E_TOL * max(1.0, abs(E_ref))
    assert _rel(grad.cpu().numpy(), g_ref) <= G_RTOL
    Eh, gh = eng.energy_grad_host(U.numpy())
    assert Eh == float(E) and np.array_equal(gh, grad.cpu().numpy())
    eng.close()


def test_general_gamma_needs_only_g_symmetry(torch_cuda):
    """A 2-RDM without any permutational symmetry: the V4 average inside the library must still
    give the exact generic gradient because g is V4-symmetric."""
    import esoo_b200
    from oracle import oracle_np as onp
    torch = torch_cuda
    M, N = 20, 6
    h, g, D, G, U = _spatial_case(torch, M, N, seed=3)
    gen = torch.Generator().manual_seed(5)
    G = torch.randn(N, N, N, N, generator=gen, dtype=torch.float64)
    D = torch.randn(N, N, generator=gen, dtype=torch.float64)
    h = torch.randn(M, M, generator=gen, dtype=torch.float64)      # non-symmetric h as well
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h, g)
    eng.set_rdms(D, G)
    E, grad = eng.energy_grad(U)
    E_ref = onp.rotated_energy_spatial(U.numpy(), D.numpy(), G.numpy(), h.numpy(), g.numpy())
    g_ref = onp.rotated_energy_grad_spatial(U.numpy(), D.numpy(), G.numpy(), h.numpy(), g.numpy())
    assert abs(float(E) - E_ref) <= E_TOL * max(1.0, abs(E_ref))
    assert _rel(grad.cpu().numpy(), g_ref) <= G_RTOL
    eng.close()


@pytest.mark.parametrize("M,N", [(8, 2), (12, 5), (20, 8), (18, 17)])
def test_generic_nonsymmetric_integrals(torch_cuda, M, N):
    """Two-body tensor, 2-RDM, h and D without any symmetry: the generic
    four-slot gradient (two dense passes) against the oracle's generic gradient."""
    import esoo_b200
    from oracle import oracle_np as onp
    from esoo_b200 import synthetic
    torch = torch_cuda
    gen = torch.Generator().manual_seed(100 + M)
    g = 0.1 * torch.randn(M, M, M, M, generator=gen, dtype=torch.float64)
    h = torch.randn(M, M, generator=gen, dtype=torch.float64)
    G = torch.randn(N, N, N, N, generator=gen, dtype=torch.float64)
    D = torch.randn(N, N, generator=gen, dtype=torch.float64)
    U = synthetic.random_partial_unitary(M, N, seed=M)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    with pytest.raises(NotImplementedError):
        eng.set_integrals(h, g, allow_generic=False)
    eng.set_integrals(h, g)
    assert eng.generic
    eng.set_rdms(D, G)
    E, grad = eng.energy_grad(U)
    E_ref = onp.rotated_energy_spatial(U.numpy(), D.numpy(), G.numpy(), h.numpy(), g.numpy())
    g_ref = onp.rotated_energy_grad_spatial(U.numpy(), D.numpy(), G.numpy(), h.numpy(), g.numpy())
    assert abs(float(E) - E_ref) <= E_TOL * max(1.0, abs(E_ref))
    assert _rel(grad.cpu().numpy(), g_ref) <= G_RTOL
    # rotated integrals do not need any symmetry either
    h_rot, g_rot = eng.transform(U)
    h_r, g_r = onp.rotated_integrals_spatial(U.numpy(), h.numpy(), g.numpy())
    assert np.max(np.abs(g_rot.cpu().numpy() - g_r)) <= 1e-11
    # the device-resident optimiser runs on the generic path as well
    res = eng.optimize(U.numpy(), 0.01, 1e-9, 25)
    hn, gn, Dn, Gn = h.numpy(), g.numpy(), D.numpy(), G.numpy()
    ref = onp.optimal_rotation(lambda X: onp.rotated_energy_spatial(X, Dn, Gn, hn, gn),
                               lambda X: onp.rotated_energy_grad_spatial(X, Dn, Gn, hn, gn),
                               U.numpy(), 0.01, 1e-9, 25)
    assert res["n_iter"] == ref["n_iter"] and abs(res["energy"] - ref["energy"]) <= EFINAL_TOL
    eng.close()


@pytest.mark.parametrize("pair_symmetric", [False, True])
@pytest.mark.parametrize("M,N,t0,mloc", [(272, 8, 268, 4), (400, 24, 100, 2), (264, 16, 0, 3),
                                         (40, 5, 11, 7), (600, 24, 37, 1)])
def test_multipass_shard_vs_oracle(torch_cuda, M, N, t0, mloc, pair_symmetric):
    """M > 256 exercises the second 256-row pass of K1 (partially filled row-blocks) on a thin
    shard of the first index, in both slab modes; the oracle contracts the same slabs.  M=400
    and M=600 with N=24 also exercise the folded partial-tile buffers (4 and 2 instead of 8)."""
    import esoo_b200
    from esoo_b200 import synthetic
    torch = torch_cuda
    h = synthetic.h_spatial(M)
    gsh = synthetic.eri_spatial_shard(M, t0, mloc, device="cuda:0")
    D, G = synthetic.rdms_spatial(N)
    U = synthetic.random_partial_unitary(M, N)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0", t0=t0, mloc=mloc)
    eng.set_integrals(h, gsh, assume_v4_symmetric=True)
    eng.set_rdms(D, G)
    eng.set_pair_symmetry(pair_symmetric)
    E, grad = eng.energy_grad(U)
    g_ref, E_ref = shard_partial_oracle(gsh.cpu().numpy(), t0, U.numpy(), D.numpy(), G.numpy(),
                                         h.numpy(), pair_symmetric)
    out = grad.cpu().numpy()
    assert abs(float(E) - E_ref) <= E_TOL * max(1.0, abs(E_ref))
    assert _rel(out, g_ref) <= G_RTOL
    if not pair_symmetric:
        mask = np.ones(M, bool)
        mask[t0:t0 + mloc] = False
        assert np.all(out[mask] == 0.0)
    eng.close()


@pytest.mark.parametrize("M,N", [(24, 4), (31, 16), (64, 24)])
def test_dense_and_pair_symmetric_modes_agree(torch_cuda, M, N):
    import esoo_b200
    from oracle import oracle_np as onp
    torch = torch_cuda
    h, g, D, G, U = _spatial_case(torch, M, N, seed=7)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h, g)
    eng.set_rdms(D, G)
    assert eng.streamed_slabs() == sum(1 for t in range(eng.M) for q in range(eng.M)
                                       if esoo_b200.distributed.pair_selected(t, q))
    E1, g1 = eng.energy_grad(U)
    eng.set_pair_symmetry(False)
    assert eng.streamed_slabs() == eng.M * eng.M
    E2, g2 = eng.energy_grad(U)
    E_ref = onp.rotated_energy_spatial(U.numpy(), D.numpy(), G.numpy(), h.numpy(), g.numpy())
    assert abs(float(E1) - float(E2)) <= 1e-12 * max(1.0, abs(E_ref))
    assert _rel(g1.cpu().numpy(), g2.cpu().numpy()) <= 1e-12
    assert abs(float(E1) - E_ref) <= E_TOL * max(1.0, abs(E_ref))
    eng.close()


@pytest.mark.parametrize("pair_symmetric", [False, True])
def test_two_shards_sum_to_full(torch_cuda, pair_symmetric):
    """World-size-2 emulation on one GPU: the shard outputs add up to the unsharded result."""
    import esoo_b200
    torch = torch_cuda
    M, N = 48, 8
    h, g, D, G, U = _spatial_case(torch, M, N, seed=1)
    full = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    full.set_integrals(h, g)
    full.set_rdms(D, G)
    E, grad = full.energy_grad(U)
    acc_E, acc_g = 0.0, torch.zeros_like(grad)
    for rank in range(2):
        t0, mloc = esoo_b200.shard_range(M, rank, 2)
        eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0", t0=t0, mloc=mloc)
        eng.set_integrals(h, g[t0:t0 + mloc], assume_v4_symmetric=True)
        eng.set_rdms(D, G)
        eng.set_pair_symmetry(pair_symmetric)
        e, gr = eng.energy_grad(U)
        acc_E += float(e)
        acc_g += gr
        eng.close()
    assert abs(acc_E - float(E)) <= 1e-11 * max(1.0, abs(float(E)))
    assert _rel(acc_g.cpu().numpy(), grad.cpu().numpy()) <= 1e-12
    full.close()


def test_transform_vs_oracle(torch_cuda):
    import esoo_b200
    from oracle import oracle_np as onp
    torch = torch_cuda
    M, N = 30, 6
    h, g, D, G, U = _spatial_case(torch, M, N, seed=2)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h, g)
    h_rot, g_rot = eng.transform(U)
    h_ref, g_ref = onp.rotated_integrals_spatial(U.numpy(), h.numpy(), g.numpy())
    assert np.max(np.abs(h_rot.cpu().numpy() - h_ref)) <= 1e-11
    assert np.max(np.abs(g_rot.cpu().numpy() - g_ref)) <= 1e-11
    eng.close()


@pytest.mark.parametrize("iteration", [0, 1, 2, 5])
def test_bb_update_vs_oracle(torch_cuda, iteration):
    import esoo_b200
    from oracle import oracle_np as onp
    torch = torch_cuda
    M, N = 26, 7
    gen = torch.Generator().manual_seed(iteration)
    Uc = torch.linalg.qr(torch.randn(M, N, generator=gen, dtype=torch.float64))[0]
    Up = torch.linalg.qr(torch.randn(M, N, generator=gen, dtype=torch.float64))[0]
    Gc = torch.randn(M, N, generator=gen, dtype=torch.float64)
    Gp = torch.randn(M, N, generator=gen, dtype=torch.float64)
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(0.07, 1e-6, 10, device="cuda:0")
    out = opt.compute_updated_partial_unitary(iteration, Uc, Up if iteration else None, Gc,
                                              Gp if iteration else None)
    ref, step = onp.bb_update(iteration, Uc.numpy(), Up.numpy(), Gc.numpy(), Gp.numpy(), 0.07)
    assert np.max(np.abs(out.cpu().numpy() - ref)) <= 1e-12
    assert abs(float(opt.BBstepsize) - step) <= 1e-13 * max(1.0, abs(step))


def test_optimize_vs_oracle_loop(torch_cuda):
    """Device-resident loop against the oracle's restatement of pupo.py:161-350 on a spatial
    problem no golden covers (N=9 -> NT=2 with padding)."""
    import esoo_b200
    from oracle import oracle_np as onp
    torch = torch_cuda
    M, N = 24, 9
    h, g, D, G, U = _spatial_case(torch, M, N, seed=4)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h, g)
    eng.set_rdms(D, G)
    res = eng.optimize(U.numpy(), 0.02, 1e-9, 250)
    hn, gn, Dn, Gn = h.numpy(), g.numpy(), D.numpy(), G.numpy()
    ref = onp.optimal_rotation(lambda X: onp.rotated_energy_spatial(X, Dn, Gn, hn, gn),
                               lambda X: onp.rotated_energy_grad_spatial(X, Dn, Gn, hn, gn),
                               U.numpy(), 0.02, 1e-9, 250)
    record_deviation("optimize_vs_oracle_loop", dE_final=abs(res["energy"] - ref["energy"]),
                     dU=np.max(np.abs(res["U"] - ref["U"])))
    assert res["n_iter"] == ref["n_iter"]
    assert abs(res["energy"] - ref["energy"]) <= EFINAL_TOL
    assert np.max(np.abs(res["U"] - ref["U"])) <= TRAJ_U_TOL
    eng.close()


def test_finite_difference_mode(torch_cuda):
    """gradient_method='finite_difference' (pupo.py:105-127): FD gradient of the CUDA energy
    agrees with the analytic CUDA gradient to FD accuracy."""
    import esoo_b200
    gold = load_golden("abba_M6_N2")
    hs, gs, Ds, Gs, U0 = golden_inputs(gold)
    fun = _Solver().compute_rotated_energy
    obj = partial(fun, oneRDM=Ds[0], twoRDM=Gs[0], one_body_integrals=hs, two_body_integrals=gs)
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 10, device="cuda:0",
                                                      gradient_method="finite_difference")
    fd = opt.compute_rotated_energy_gradient(U0.clone(), obj).cpu().numpy()
    assert _rel(fd, gold["grad"]) <= 1e-6
    esoo_b200.clear_engine_cache()


def test_rejects_unknown_objective_and_complex(torch_cuda):
    import esoo_b200
    torch = torch_cuda
    gold = load_golden("abba_M6_N2")
    hs, gs, Ds, Gs, U0 = golden_inputs(gold)
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 10, device="cuda:0")
    with pytest.raises(TypeError):
        opt.compute_optimal_rotation(lambda **k: 0.0, U0, Ds[0], Gs[0], hs, gs)
    with pytest.raises(NotImplementedError):
        opt.compute_optimal_rotation(_Solver().compute_rotated_energy, U0,
                                     Ds[0].to(torch.complex128), Gs[0].to(torch.complex128), hs, gs)
    esoo_b200.clear_engine_cache()


# --------------------------------------------------------------------------------------------
# (c) BASELINE.json size: M=256, N=16 — properties that do not need a CPU answer
# --------------------------------------------------------------------------------------------
def test_full_size_properties(torch_cuda):
    import esoo_b200
    from esoo_b200 import synthetic
    torch = torch_cuda
    M, N = 256, 16
    free, _ = torch.cuda.mem_get_info(0)
    if free < 40 * (1 << 30):
        pytest.skip("needs 40 GB of free device memory")
    h = synthetic.h_spatial(M).cuda()
    g = synthetic.eri_spatial(M, device="cuda:0")
    U = synthetic.random_partial_unitary(M, N).cuda()
    D1, G1 = synthetic.rdms_spatial(N, seed=11)
    D2, G2 = synthetic.rdms_spatial(N, seed=12)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h, g)                      # includes the on-device V4 symmetry verification
    a, b = 0.7, -1.3
    eng.set_rdms(D1, G1)
    E1, g1 = eng.energy_grad(U)
    eng.set_rdms(D2, G2)
    E2, g2 = eng.energy_grad(U)
    eng.set_rdms(a * D1 + b * D2, a * G1 + b * G2)
    E3, g3 = eng.energy_grad(U)
    # linearity in the RDMs (eig.py:149-169 relies on it)
    scale = max(abs(float(E1)), abs(float(E2)), 1.0)
    assert abs(float(E3) - (a * float(E1) + b * float(E2))) <= 1e-10 * scale
    assert _rel(g3.cpu().numpy(), (a * g1 + b * g2).cpu().numpy()) <= 1e-11
    # gradient = directional derivative of the energy (central difference along a random X)
    gen = torch.Generator().manual_seed(3)
    X = torch.randn(M, N, generator=gen, dtype=torch.float64).cuda()
    X /= X.norm()
    eps = 1e-4
    Ep, _ = eng.energy_grad(U + eps * X)
    Em, _ = eng.energy_grad(U - eps * X)
    fd = (float(Ep) - float(Em)) / (2 * eps)
    an = float((g3 * X).sum())
    assert abs(fd - an) <= 1e-6 * max(1.0, abs(an))
    # determinism: bit-identical repeat
    E4, g4 = eng.energy_grad(U)
    assert float(E4) == float(E3) and torch.equal(g4, g3)
    # shard additivity at full size (8 shards, as on an 8-GPU box)
    accE, accg = 0.0, torch.zeros_like(g3)
    for r in range(8):
        t0, mloc = esoo_b200.shard_range(M, r, 8)
        sh = esoo_b200.OrbitalEngine(M, N, device="cuda:0", t0=t0, mloc=mloc)
        sh.set_integrals(h, g[t0:t0 + mloc], assume_v4_symmetric=True)
        sh.set_rdms(a * D1 + b * D2, a * G1 + b * G2)
        e, gr = sh.energy_grad(U)
        accE += float(e)
        accg += gr
        sh.close()
    assert abs(accE - float(E3)) <= 1e-10 * scale
    assert _rel(accg.cpu().numpy(), g3.cpu().numpy()) <= 1e-12
    eng.close()


# --------------------------------------------------------------------------------------------
# (d) the reference's outer loop (exact-diagonalisation stand-in for VQE):
#     goldens produced with the LIVE reference optimiser on the CPU
# --------------------------------------------------------------------------------------------
from conftest import outer_golden_names  # noqa: E402


@pytest.mark.parametrize("name", outer_golden_names())
def test_outer_loop_vs_reference_golden(torch_cuda, name):
    """Final (and every intermediate) outer-loop energy within 1e-8 Ha of the run that used the
    reference optimiser; the rotated Hamiltonian comes from the CUDA transform (oo_transform)."""
    import esoo_b200
    from esoo_b200 import harness, ingest
    torch = torch_cuda
    gold = load_golden(name)
    hs, gs = torch.from_numpy(gold["h_spin"]), torch.from_numpy(gold["g_spin"])
    M, N, k = int(gold["M"]), int(gold["N"]), int(gold["n_states"])
    weights = list(gold["weights"]) if k > 1 else None
    h_sp, g_sp, _ = ingest.reduce_integrals(hs, gs)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h_sp, g_sp)
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(float(gold["bb0"]), float(gold["tol"]),
                                                      int(gold["maxiter"]), device="cuda:0")
    res = harness.run_outer_loop(opt, hs, gs, 2 * N, int(gold["n_alpha"]), int(gold["n_beta"]),
                                 maxiter=int(gold["outer_maxiter"]),
                                 stopping_tolerance=float(gold["outer_tol"]), n_states=k,
                                 weights=weights, engine_for_transform=eng)
    E = np.array(res["energies"])
    assert E.shape == gold["energies"].shape, "different number of outer iterations"
    assert np.max(np.abs(E - gold["energies"])) <= EFINAL_TOL
    if "ref_test_golden" in gold:
        # H2 / 6-31G: the number hard-coded in the reference's own tests (test_optorbvqe.py:67,
        # test_optorbmcvqe.py:61; their bar is decimal=3)
        assert np.max(np.abs(E[-1] - gold["ref_test_golden"])) < 2e-5
    eng.close()
    esoo_b200.clear_engine_cache()


@pytest.mark.parametrize("pattern", ["abba", "abab"])
def test_device_ingest_matches_torch_ingest(torch_cuda, pattern):
    """Spin-orbital -> spatial reduction done by the library's kernels (oo_ingest_spin_g,
    oo_set_rdms_spin) against the torch host logic (which the CPU suite pins to the reference)."""
    import esoo_b200
    from esoo_b200 import ingest, synthetic
    torch = torch_cuda
    M, N = 7, 3
    h, g = synthetic.h_spatial(M), synthetic.eri_spatial(M)
    hs, gs = synthetic.spin_orbital_integrals(h, g, pattern)
    h1, g1, st1 = ingest.reduce_integrals(hs, gs)
    h2, g2, st2 = ingest.reduce_integrals_device(hs.cuda(), gs.cuda())
    assert sorted(st1.blocks) == sorted(st2.blocks)
    assert torch.equal(h1, h2.cpu()) and torch.equal(g1, g2.cpu())
    bad = gs.clone()
    bad[:M, M:, M:, :M] *= 1.5                       # break the equality of the non-zero blocks
    bad[:M, M:, :M, M:] *= 1.5
    with pytest.raises(NotImplementedError):
        ingest.reduce_integrals_device(hs.cuda(), bad.cuda())
    # RDMs: 3 weighted states through both routes must give the same energy and gradient
    Ds, Gs = zip(*[synthetic.rdms_spin(N, seed=50 + n) for n in range(3)])
    w = [3.0, 2.0, 1.0]
    D_sp, G_sp = ingest.reduce_rdms(list(Ds), list(Gs), st1, w)
    U = synthetic.random_partial_unitary(M, N)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h1, g1)
    eng.set_rdms(D_sp, G_sp)
    E1, g_1 = eng.energy_grad(U)
    eng.set_rdms_spin(list(Ds), list(Gs), w, ingest.block_mask(st2))
    E2, g_2 = eng.energy_grad(U)
    assert abs(float(E1) - float(E2)) <= 1e-12 * max(1.0, abs(float(E1)))
    assert _rel(g_2.cpu().numpy(), g_1.cpu().numpy()) <= 1e-12
    eng.close()


@pytest.mark.parametrize("maxiter", [0, 1, 2, 3, 4, 7])
def test_optimizer_small_maxiter_matches_oracle(torch_cuda, maxiter):
    """The three hand-unrolled iterations always run and the loop test is `k <= maxiter`
    (pupo.py:191-304): same iteration count, energy and U as the oracle for tiny maxiter."""
    import esoo_b200
    from oracle import oracle_np as onp
    torch = torch_cuda
    M, N = 14, 3
    h, g, D, G, U = _spatial_case(torch, M, N, seed=9)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h, g)
    eng.set_rdms(D, G)
    res = eng.optimize(U.numpy(), 0.02, 1e-14, maxiter)
    hn, gn, Dn, Gn = h.numpy(), g.numpy(), D.numpy(), G.numpy()
    ref = onp.optimal_rotation(lambda X: onp.rotated_energy_spatial(X, Dn, Gn, hn, gn),
                               lambda X: onp.rotated_energy_grad_spatial(X, Dn, Gn, hn, gn),
                               U.numpy(), 0.02, 1e-14, maxiter)
    assert res["n_iter"] == ref["n_iter"] == max(3, maxiter + 1)
    assert abs(res["energy"] - ref["energy"]) <= 1e-10
    assert np.max(np.abs(res["U"] - ref["U"])) <= 1e-9
    assert abs(res["stepsize"] - ref["stepsize"]) <= 1e-9 * max(1.0, abs(ref["stepsize"]))
    eng.close()


def test_call_sequence_errors(torch_cuda):
    """The C layer reports misuse through status codes + oo_last_error (no abort, no fallback)."""
    import esoo_b200
    from esoo_b200._lib import OOError
    torch = torch_cuda
    M, N = 8, 2
    h, g, D, G, U = _spatial_case(torch, M, N)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    with pytest.raises(OOError, match="oo_set_integrals"):
        eng.energy_grad(U)
    eng.set_integrals(h, g)
    with pytest.raises(OOError, match="oo_set_rdms"):
        eng.energy_grad(U)
    with pytest.raises(ValueError):
        eng.set_rdms(D[:1], G)
    with pytest.raises(TypeError):
        eng.set_rdms(D.float(), G.float())
    eng.set_rdms(D, G)
    E, _ = eng.energy_grad(U)
    assert np.isfinite(float(E))
    eng.close()
    with pytest.raises(OOError):
        esoo_b200.OrbitalEngine(4, 8, device="cuda:0")          # N > M
    with pytest.raises(OOError):
        esoo_b200.OrbitalEngine(80, 40, device="cuda:0")        # N > 32


def test_jacobi_fallback_path(torch_cuda, monkeypatch):
    """The retraction's eigensolver path (taken when Newton-Schulz does not converge) forced via
    OO_FORCE_JACOBI: orth() against the reference golden and the optimiser against the oracle."""
    import esoo_b200
    from oracle import oracle_np as onp
    torch = torch_cuda
    monkeypatch.setenv("OO_FORCE_JACOBI", "1")
    gold = load_golden("abba_M12_N4")
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 10, device="cuda:0")
    out = opt.orth(torch.from_numpy(gold["V"])).cpu().numpy()
    assert np.max(np.abs(out - gold["orthV"])) <= 1e-12
    M, N = 24, 9
    h, g, D, G, U = _spatial_case(torch, M, N, seed=4)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h, g)
    eng.set_rdms(D, G)
    res = eng.optimize(U.numpy(), 0.02, 1e-9, 60)
    assert res["jacobi_fallbacks"] == res["n_iter"] and res["newton_schulz_iterations"] == 0
    hn, gn, Dn, Gn = h.numpy(), g.numpy(), D.numpy(), G.numpy()
    ref = onp.optimal_rotation(lambda X: onp.rotated_energy_spatial(X, Dn, Gn, hn, gn),
                               lambda X: onp.rotated_energy_grad_spatial(X, Dn, Gn, hn, gn),
                               U.numpy(), 0.02, 1e-9, 60)
    assert res["n_iter"] == ref["n_iter"] and abs(res["energy"] - ref["energy"]) <= EFINAL_TOL
    eng.close()


# --------------------------------------------------------------------------------------------
# pair-packed storage (OO_G_PAIR_PACKED): half the memory, same streamed slabs
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,t0,mloc", [(24, 4, 0, 24), (72, 16, 0, 72), (300, 8, 290, 6),
                                         (400, 24, 100, 2), (40, 5, 11, 7)])
def test_pair_packed_storage_is_bit_identical(torch_cuda, M, N, t0, mloc):
    """Packed and dense storage stream the same slabs in the same order, so energy and gradient
    agree bit for bit; the device gather (oo_pack_pair_slabs) and the slab-wise synthetic
    generator produce the same packed tensor."""
    import esoo_b200
    from esoo_b200 import synthetic
    torch = torch_cuda
    h = synthetic.h_spatial(M)
    gsh = synthetic.eri_spatial_shard(M, t0, mloc, device="cuda:0")
    D, G = synthetic.rdms_spatial(N)
    U = synthetic.random_partial_unitary(M, N)
    dense = esoo_b200.OrbitalEngine(M, N, device="cuda:0", t0=t0, mloc=mloc)
    dense.set_integrals(h, gsh, assume_v4_symmetric=True)
    dense.set_rdms(D, G)
    E0, g0 = dense.energy_grad(U)
    packed_t = dense.pack_pair_slabs(gsh)
    assert packed_t.shape[0] == dense.streamed_slabs() == dense.streamed_slabs_pair()
    assert torch.equal(packed_t, synthetic.eri_spatial_pair_packed(M, t0, mloc, device="cuda:0"))
    lst = esoo_b200.distributed.pair_slab_list(M, t0, mloc)
    for i in (0, len(lst) // 2, len(lst) - 1):
        t, q = lst[i]
        assert torch.equal(packed_t[i], gsh[t - t0, q])
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0", t0=t0, mloc=mloc)
    eng.set_integrals_packed(h, packed_t)
    eng.set_rdms(D, G)
    E1, g1 = eng.energy_grad(U)
    assert float(E1) == float(E0)
    assert torch.equal(g1, g0)
    with pytest.raises(RuntimeError):
        eng.set_pair_symmetry(False)              # the other half of the slabs is not stored
    if mloc == M:
        h0, gr0 = dense.transform(U)
        h1, gr1 = eng.transform(U)
        assert torch.equal(h0, h1) and torch.equal(gr0, gr1)
        r0 = dense.optimize(U.numpy(), 0.02, 1e-9, 40)
        r1 = eng.optimize(U.numpy(), 0.02, 1e-9, 40)
        assert r0["n_iter"] == r1["n_iter"] and r0["energy"] == r1["energy"]
        assert np.array_equal(r0["U"], r1["U"])
        # back to dense storage on the same context
        eng.set_integrals(h, gsh, assume_v4_symmetric=True)
        E2, g2 = eng.energy_grad(U)
        assert float(E2) == float(E0) and torch.equal(g2, g0)
    eng.close()
    dense.close()


def test_inputs_on_host_option(torch_cuda):
    """inputs_on_host=True: `.device` reads 'cpu', so the outer loop hands over host tensors (no
    per-iteration H2D of the (2M)^4 tensor); the result is the same as with device inputs."""
    import esoo_b200
    from esoo_b200 import harness
    torch = torch_cuda
    gold = load_golden("outer_H4_631G_ground")
    hs, gs = torch.from_numpy(gold["h_spin"]), torch.from_numpy(gold["g_spin"])
    N = int(gold["N"])
    runs = []
    for on_host in (False, True):
        esoo_b200.clear_engine_cache()
        opt = esoo_b200.PartialUnitaryProjectionOptimizer(float(gold["bb0"]), float(gold["tol"]),
                                                          int(gold["maxiter"]), device="cuda:0",
                                                          inputs_on_host=on_host)
        assert opt.device == ("cpu" if on_host else "cuda:0") and opt.compute_device == "cuda:0"
        res = harness.run_outer_loop(opt, hs, gs, 2 * N, int(gold["n_alpha"]), int(gold["n_beta"]),
                                     maxiter=int(gold["outer_maxiter"]),
                                     stopping_tolerance=float(gold["outer_tol"]))
        runs.append(np.array(res["energies"]))
        assert res["U"].device.type == "cpu"
    assert np.array_equal(runs[0], runs[1])
    assert np.max(np.abs(runs[1] - gold["energies"])) <= EFINAL_TOL
    esoo_b200.clear_engine_cache()


def test_bb_update_vs_reference_golden(torch_cuda):
    """oo_bb_update through the drop-in method against the live reference's
    compute_updated_partial_unitary (fixture bb_update_M14_N3)."""
    import esoo_b200
    torch = torch_cuda
    gold = load_golden("bb_update_M14_N3")
    t = {k: torch.from_numpy(gold[k]) for k in ("U_cur", "U_prev", "G_cur", "G_prev")}
    for it in gold["iterations"]:
        opt = esoo_b200.PartialUnitaryProjectionOptimizer(float(gold["bb0"]), 1e-6, 10,
                                                          device="cuda:0")
        out = opt.compute_updated_partial_unitary(int(it), t["U_cur"], t["U_prev"], t["G_cur"],
                                                  t["G_prev"])
        assert np.max(np.abs(out.cpu().numpy() - gold[f"U_next_{it}"])) <= 1e-12
        assert abs(float(opt.BBstepsize) - float(gold[f"step_{it}"])) <= \
            1e-13 * abs(float(gold[f"step_{it}"]))


def test_decay_factor_vs_reference_golden(torch_cuda):
    """decay_factor reaches the device-side stopping rule: same callback iterations and final
    energy as the live reference for 0.2 / 0.5 / 0.95."""
    import esoo_b200
    gold = load_golden("opt_decay_M6_N2")
    hs, gs, Ds, Gs, U0 = golden_inputs(gold)
    for d in gold["decays"]:
        calls = []
        opt = esoo_b200.PartialUnitaryProjectionOptimizer(
            float(gold["bb0"]), float(gold["tol"]), int(gold["maxiter"]),
            callback=lambda it, e: calls.append(it), decay_factor=float(d), device="cuda:0")
        U, E = opt.compute_optimal_rotation(fun=_Solver().compute_rotated_energy,
                                            initial_partial_unitary=U0.clone(), oneRDM=Ds[0],
                                            twoRDM=Gs[0], one_body_integrals=hs,
                                            two_body_integrals=gs)
        record_deviation(f"decay:{d}", dE_final=abs(float(E) - float(gold[f"E_{d}"])),
                         dU=np.max(np.abs(U.numpy() - gold[f"U_{d}"])))
        assert calls == list(gold[f"calls_it_{d}"])
        assert abs(float(E) - float(gold[f"E_{d}"])) <= EFINAL_TOL
        assert np.max(np.abs(U.numpy() - gold[f"U_{d}"])) <= TRAJ_U_TOL
    esoo_b200.clear_engine_cache()


# --------------------------------------------------------------------------------------------
# round-2 additions
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rank", range(8))
def test_headline_shard_vs_oracle(torch_cuda, rank):
    """The exact shard one GPU of the 8-GPU BASELINE run evaluates (M=256, N=16, 32 rows of the
    first index, pair-symmetric mode), slab for slab against the numpy oracle."""
    import esoo_b200
    from esoo_b200 import synthetic
    M, N = 256, 16
    t0, mloc = esoo_b200.shard_range(M, rank, 8)
    assert mloc == 32
    h = synthetic.h_spatial(M)
    gsh = synthetic.eri_spatial_shard(M, t0, mloc, device="cuda:0")
    D, G = synthetic.rdms_spatial(N)
    U = synthetic.random_partial_unitary(M, N)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0", t0=t0, mloc=mloc)
    eng.set_integrals(h, gsh, assume_v4_symmetric=True)
    eng.set_rdms(D, G)
    E, grad = eng.energy_grad(U)
    g_ref, E_ref = shard_partial_oracle(gsh.cpu().numpy(), t0, U.numpy(), D.numpy(), G.numpy(),
                                        h.numpy(), True)
    dE = abs(float(E) - E_ref) / max(1.0, abs(E_ref))
    dg = _rel(grad.cpu().numpy(), g_ref)
    record_deviation(f"headline_shard:{rank}", dE_rel=dE, dgrad_rel=dg)
    assert dE <= E_TOL and dg <= G_RTOL
    eng.close()


def test_finite_difference_optimal_rotation_vs_reference_golden(torch_cuda):
    """gradient_method='finite_difference' through compute_optimal_rotation (pupo.py:105-127,
    183-184) against the live reference's FD run: same callback iterations; FD gradients carry
    ~1e-8 relative noise, so energies agree to ~1e-6 rather than to machine precision."""
    import esoo_b200
    torch = torch_cuda
    gold = load_golden("fd_M5_N2")
    hs, gs = torch.from_numpy(gold["h_spin"]), torch.from_numpy(gold["g_spin"])
    D, G = torch.from_numpy(gold["D_spin_0"]), torch.from_numpy(gold["G_spin_0"])
    calls = []
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(
        float(gold["bb0"]), float(gold["tol"]), int(gold["maxiter"]),
        callback=lambda it, e: calls.append((it, e)), gradient_method="finite_difference",
        device="cuda:0")
    obj = partial(_Solver().compute_rotated_energy, oneRDM=D, twoRDM=G, one_body_integrals=hs,
                  two_body_integrals=gs)
    fd0 = opt.compute_rotated_energy_gradient(torch.from_numpy(gold["U0"]), obj).cpu().numpy()
    assert np.max(np.abs(fd0 - gold["fd_grad_U0"])) <= 1e-6
    U, E = opt.compute_optimal_rotation(fun=_Solver().compute_rotated_energy,
                                        initial_partial_unitary=torch.from_numpy(gold["U0"]),
                                        oneRDM=D, twoRDM=G, one_body_integrals=hs,
                                        two_body_integrals=gs)
    dcb = np.max(np.abs(np.array([c[1] for c in calls]) - gold["opt_calls_E"]))
    record_deviation("fd_trajectory", dE_final=abs(float(E) - float(gold["opt_E"])), dE_callbacks=dcb,
                     dU=np.max(np.abs(U.numpy() - gold["opt_U"])))
    assert [c[0] for c in calls] == list(gold["opt_calls_it"])
    # measured on the B200: callbacks 8.1e-8, final energy 1.5e-9, final U 9.3e-9
    assert dcb <= 1e-6 and abs(float(E) - float(gold["opt_E"])) <= 1e-7
    assert opt.last_result["n_iter"] == int(gold["opt_calls_it"][-1]) + 1
    esoo_b200.clear_engine_cache()


def test_rotated_hamiltonian_binding_vs_reference_golden(torch_cuda):
    """h', g' of get_rotated_hamiltonian (base_opt_orb_solver.py:597-604) through the binding
    (oo_transform on the optimiser's cached engine + spin-block embedding) against the
    reference's own einsums, both spin patterns, odd M."""
    import esoo_b200
    torch = torch_cuda
    gold = load_golden("rotated_integrals")
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 10, device="cuda:0")
    for tag in "abc":
        hs, gs = torch.from_numpy(gold[f"{tag}_h_spin"]), torch.from_numpy(gold[f"{tag}_g_spin"])
        U = torch.from_numpy(gold[f"{tag}_U"])
        h_rot, g_rot = esoo_b200.rotated_spin_integrals(opt, hs, gs, U)
        assert h_rot.shape == gold[f"{tag}_h_rot"].shape and g_rot.shape == gold[f"{tag}_g_rot"].shape
        assert np.max(np.abs(h_rot - gold[f"{tag}_h_rot"])) <= 1e-12
        assert np.max(np.abs(g_rot - gold[f"{tag}_g_rot"])) <= 1e-12

    class Solver(esoo_b200.RotatedHamiltonianMixin):          # what a reference subclass provides
        def __init__(self, h, g, optimizer):
            self.one_body_integrals, self.two_body_integrals = h, g
            self._partial_unitary_optimizer_list = [None, optimizer]

    sv = Solver(torch.from_numpy(gold["a_h_spin"]), torch.from_numpy(gold["a_g_spin"]), opt)
    h_rot, g_rot = sv.rotated_integral_tensors(torch.from_numpy(gold["a_U"]))
    assert np.max(np.abs(g_rot - gold["a_g_rot"])) <= 1e-12
    esoo_b200.clear_engine_cache()


def test_engine_cache_detects_changed_integrals(torch_cuda):
    """a tensor that differs in ONE element (in place, or a modified copy) must not be
    served by the engine cached for the original; the same objects, unmodified, are a hit."""
    import esoo_b200
    from esoo_b200 import optimizer as om
    torch = torch_cuda
    gold = load_golden("abba_M6_N2")
    hs, gs, Ds, Gs, U0 = golden_inputs(gold)
    fun = _Solver().compute_rotated_energy
    esoo_b200.clear_engine_cache()
    for where in ("cpu", "cuda"):
        h_in, g_in = hs.to(where), gs.clone().to(where)
        opt = esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 10, device="cuda:0")
        eng0 = opt._prepare(fun, Ds[0], Gs[0], h_in, g_in, 2)
        E0 = eng0.energy_grad_host(U0.numpy())[0]
        assert opt._prepare(fun, Ds[0], Gs[0], h_in, g_in, 2) is eng0       # same objects: hit
        assert opt._prepare(fun, Ds[0], Gs[0], h_in.clone(), g_in.clone(), 2) is eng0  # same content
        g_mod = g_in.clone()
        for p_, q_, r_, s_ in ((1, 2, 3, 2), (1, 8, 9, 2), (7, 2, 3, 8), (7, 8, 9, 8)):
            g_mod[p_, q_, r_, s_] += 1e-3       # spatial element (1,2,3,2) in all four spin blocks
        eng1 = opt._prepare(fun, Ds[0], Gs[0], h_in, g_mod, 2)
        assert eng1 is not eng0 and eng1.generic     # a new engine (the edit also broke V4 symmetry)
        assert eng1.energy_grad_host(U0.numpy())[0] != E0
        # in-place edit of the very tensor the engine was built from (version counter moves)
        g_in[0, 6, 6, 0] *= 1.0 + 1e-9
        g_in[6, 0, 0, 6] *= 1.0 + 1e-9
        g_in[0, 0, 0, 0] *= 1.0 + 1e-9
        g_in[6, 6, 6, 6] *= 1.0 + 1e-9
        eng2 = opt._prepare(fun, Ds[0], Gs[0], h_in, g_in, 2)
        assert eng2 is not eng0
        assert eng2.energy_grad_host(U0.numpy())[0] != E0
        esoo_b200.clear_engine_cache()
    a = torch.randn(5, 7, 3, dtype=torch.float64)
    b = a.clone()
    b[2, 3, 1] += 1e-9
    assert om.content_checksum(a) != om.content_checksum(b)
    assert om.content_checksum(a) != om.content_checksum(a.flip(0))
    assert om.content_checksum(a) == om.content_checksum(a.clone())


def test_callback_exception_propagates(torch_cuda):
    """An exception raised by the user's callback leaves compute_optimal_rotation (as in the
    reference, pupo.py:193-194) instead of being swallowed by ctypes; the device loop stops and
    the engine stays usable."""
    import esoo_b200
    gold = load_golden("abba_M8_N3")
    hs, gs, Ds, Gs, U0 = golden_inputs(gold)

    class Abort(Exception):
        pass

    seen = []

    def cb(it, e):
        seen.append(it)
        if it == 9:
            raise Abort("stop here")

    opt = esoo_b200.PartialUnitaryProjectionOptimizer(float(gold["opt_bb0"]), float(gold["opt_tol"]),
                                                      int(gold["opt_maxiter"]), callback=cb,
                                                      device="cuda:0")
    with pytest.raises(Abort):
        opt.compute_optimal_rotation(fun=_Solver().compute_rotated_energy,
                                     initial_partial_unitary=U0.clone(), oneRDM=Ds[0], twoRDM=Gs[0],
                                     one_body_integrals=hs, two_body_integrals=gs)
    assert seen == list(range(10))
    opt.callback = None
    U, E = opt.compute_optimal_rotation(fun=_Solver().compute_rotated_energy,
                                        initial_partial_unitary=U0.clone(), oneRDM=Ds[0],
                                        twoRDM=Gs[0], one_body_integrals=hs, two_body_integrals=gs)
    assert abs(float(E) - float(gold["opt_E"])) <= EFINAL_TOL
    esoo_b200.clear_engine_cache()


def test_pipelined_host_evaluations(torch_cuda):
    """oo_eval_submit / oo_eval_wait (two slots in flight) return exactly what the synchronous
    host-buffer call returns, in order; a slot cannot be submitted twice."""
    import esoo_b200
    from esoo_b200 import synthetic
    from esoo_b200._lib import OOError
    torch = torch_cuda
    M, N = 37, 6                                         # odd M: padded inside the engine
    h, g, D, G, U = _spatial_case(torch, M, N, seed=5)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h, g)
    eng.set_rdms(D, G)
    Us = [synthetic.random_partial_unitary(M, N, seed=100 + i).numpy() for i in range(7)]
    ref = [eng.energy_grad_host(u) for u in Us]
    assert np.array_equal(eng.energies_host(Us), np.array([r[0] for r in ref]))
    eng.submit_host(Us[0], 0)
    eng.submit_host(Us[1], 1)
    with pytest.raises(OOError):
        eng.submit_host(Us[2], 0)
    for s in (0, 1):
        E, grad = eng.wait_host(s)
        assert E == ref[s][0] and np.array_equal(grad, ref[s][1])
    with pytest.raises(OOError):
        eng.wait_host(0)
    eng.close()


def test_spatial_integrals_through_the_class(torch_cuda):
    """SpatialIntegrals (extension input format for problems whose (2M)^4 tensor cannot exist):
    dense, pair-packed and non-symmetric spatial tensors through compute_optimal_rotation give the
    result of the spin-orbital route."""
    import esoo_b200
    from esoo_b200 import synthetic
    torch = torch_cuda
    gold = load_golden("abba_M12_N4")
    hs, gs, Ds, Gs, U0 = golden_inputs(gold)
    M = 12
    h_sp, g_sp = hs[:M, :M].contiguous(), gs[:M, M:, M:, :M].contiguous()
    fun = _Solver().compute_rotated_energy
    results = []
    eng = esoo_b200.OrbitalEngine(M, 4, device="cuda:0")
    packed = eng.pack_pair_slabs(g_sp.cuda())
    eng.close()
    for sp in (esoo_b200.SpatialIntegrals(g_sp.cuda(), M),
               esoo_b200.SpatialIntegrals(packed, M, packed=True),
               esoo_b200.SpatialIntegrals(g_sp.cuda(), M, v4_symmetric=False,
                                          g_pair_transposed=g_sp.permute(2, 3, 0, 1).contiguous().cuda())):
        opt = esoo_b200.PartialUnitaryProjectionOptimizer(
            float(gold["opt_bb0"]), float(gold["opt_tol"]), int(gold["opt_maxiter"]), device="cuda:0")
        U, E = opt.compute_optimal_rotation(fun=fun, initial_partial_unitary=U0.clone(),
                                            oneRDM=Ds[0], twoRDM=Gs[0], one_body_integrals=h_sp,
                                            two_body_integrals=sp)
        assert abs(float(E) - float(gold["opt_E"])) <= EFINAL_TOL
        results.append(U.numpy())
        esoo_b200.clear_engine_cache()
    assert np.array_equal(results[0], results[1])          # packed = dense, bit for bit
    record_deviation("spatial_generic_vs_v4", dU=np.max(np.abs(results[2] - results[0])))
    assert np.max(np.abs(results[2] - results[0])) <= 1e-7


def test_generic_two_shards_sum_to_full(torch_cuda):
    """Tensors without any symmetry on two first-index shards (each with its rows of the
    pair-transposed tensor): the partial (gradient | energy) buffers add up to the oracle."""
    import esoo_b200
    from esoo_b200 import synthetic
    from oracle import oracle_np as onp
    torch = torch_cuda
    M, N = 14, 5
    gen = torch.Generator().manual_seed(77)
    g = 0.1 * torch.randn(M, M, M, M, generator=gen, dtype=torch.float64)
    h = torch.randn(M, M, generator=gen, dtype=torch.float64)
    G = torch.randn(N, N, N, N, generator=gen, dtype=torch.float64)
    D = torch.randn(N, N, generator=gen, dtype=torch.float64)
    U = synthetic.random_partial_unitary(M, N, seed=3)
    g_pt = g.permute(2, 3, 0, 1).contiguous()
    accE, accg = 0.0, np.zeros((M, N))
    for rank in range(2):
        t0, mloc = esoo_b200.shard_range(M, rank, 2)
        eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0", t0=t0, mloc=mloc)
        eng.set_integrals(h, g[t0:t0 + mloc], g_pair_transposed=g_pt[t0:t0 + mloc])
        assert eng.generic
        eng.set_rdms(D, G)
        e, gr = eng.energy_grad(U)
        accE += float(e)
        accg += gr.cpu().numpy()
        eng.close()
    E_ref = onp.rotated_energy_spatial(U.numpy(), D.numpy(), G.numpy(), h.numpy(), g.numpy())
    g_ref = onp.rotated_energy_grad_spatial(U.numpy(), D.numpy(), G.numpy(), h.numpy(), g.numpy())
    assert abs(accE - E_ref) <= E_TOL * max(1.0, abs(E_ref))
    assert _rel(accg, g_ref) <= G_RTOL


def test_step_fusion_matches_separate_step(torch_cuda, monkeypatch):
    """The optimiser transition fused into the tail kernel's last CTA (256 threads) and the
    stand-alone k_step launch (1024 threads) walk the same trajectory."""
    import esoo_b200
    torch = torch_cuda
    M, N = 30, 7
    h, g, D, G, U = _spatial_case(torch, M, N, seed=8)
    runs = []
    for fused in (True, False):
        if not fused:
            monkeypatch.setenv("OO_NO_STEP_FUSION", "1")
        eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
        eng.set_integrals(h, g)
        eng.set_rdms(D, G)
        runs.append(eng.optimize(U.numpy(), 0.02, 1e-9, 80))
        eng.close()
    assert runs[0]["n_iter"] == runs[1]["n_iter"]
    assert abs(runs[0]["energy"] - runs[1]["energy"]) <= 1e-10
    assert np.max(np.abs(runs[0]["U"] - runs[1]["U"])) <= 1e-8


@pytest.mark.parametrize("M,N", [(30, 26), (40, 32)])
def test_optimize_on_the_tiles_path_vs_oracle(torch_cuda, M, N):
    """N in 25..32 keeps the stored-tiles evaluation (k_qcontract + k_tail_row);
    the optimiser transition is fused into k_tail_row's last CTA there: same trajectory as the
    oracle's restatement of pupo.py:161-350."""
    import esoo_b200
    from oracle import oracle_np as onp
    torch = torch_cuda
    h, g, D, G, U = _spatial_case(torch, M, N, seed=M)
    eng = esoo_b200.OrbitalEngine(M, N, device="cuda:0")
    eng.set_integrals(h, g)
    eng.set_rdms(D, G)
    res = eng.optimize(U.numpy(), 0.01, 1e-9, 30)
    hn, gn, Dn, Gn = h.numpy(), g.numpy(), D.numpy(), G.numpy()
    ref = onp.optimal_rotation(lambda X: onp.rotated_energy_spatial(X, Dn, Gn, hn, gn),
                               lambda X: onp.rotated_energy_grad_spatial(X, Dn, Gn, hn, gn),
                               U.numpy(), 0.01, 1e-9, 30)
    record_deviation(f"tiles_path_optimize:{M}x{N}", dE_final=abs(res["energy"] - ref["energy"]),
                     dU=np.max(np.abs(res["U"] - ref["U"])))
    assert res["n_iter"] == ref["n_iter"]
    assert abs(res["energy"] - ref["energy"]) <= EFINAL_TOL * max(1.0, abs(ref["energy"]))
    assert np.max(np.abs(res["U"] - ref["U"])) <= 1e-8
    eng.close()
