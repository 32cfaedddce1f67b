"""Multi-GPU parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/multigpu_check.py

Every rank holds a shard of g (first index).  Checked against the numpy ORACLE (tests-only CPU
restatement of the reference, oracle/oracle_np.py) and, for the larger shapes, the unsharded engine:
  1. all-reduced E and dE/dU, both all-reduce implementations (NCCL; fused into the tail kernel
     over NVLink peer memory), pair-symmetric / dense / pair-packed / generic (no symmetry) shards;
  2. the device-resident optimiser on the sharded tensor: iteration count, final energy and U
     against the oracle's restatement of pupo.py:161-350;
  3. the drop-in class itself under torch.distributed: PartialUnitaryProjectionOptimizer shards
     the reference's spin-orbital tensors behind compute_optimal_rotation -- a complete outer loop
     (fixture outer_H4_631G_ground, produced with the live reference optimiser) to 1e-8 Ha, the
     rotated-Hamiltonian binding, and SpatialIntegrals shards;
  4. (OO_MG_CONFIG5=1) BASELINE config 5, M=400, N=24, through the class on all GPUs of the box;
  5. a rank left alone in the fused all-reduce fails with an error and a NaN result (time-out).
Prints MULTIGPU PASS.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import esoo_b200  # noqa: E402
from esoo_b200 import harness, synthetic  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402

E_TOL, G_RTOL, EFINAL_TOL = 1e-10, 1e-9, 1e-8


def log(rank, msg):
    if rank == 0:
        print(msg, flush=True)


def engine_cases(rank, world, dev):
    ok = True
    # (M, N, mode): "pair" / "dense" / "packed" / "generic"
    cases = [(64, 16, "pair"), (64, 16, "dense"), (50, 5, "pair"), (64, 16, "packed"),
             (22, 5, "generic"), (272, 8, "pair")]
    cases = cases[:int(os.environ.get("OO_MG_CASES", len(cases)))]
    for (M, N, mode) in cases:
        if world > M:
            continue
        D, G = synthetic.rdms_spatial(N)
        h = synthetic.h_spatial(M)
        U = synthetic.random_partial_unitary(M, N)
        t0, mloc = esoo_b200.shard_range(M, rank, world)
        if mode == "generic":
            gen = torch.Generator().manual_seed(5)
            g_full = 0.1 * torch.randn(M, M, M, M, generator=gen, dtype=torch.float64)
            G = torch.randn(N, N, N, N, generator=gen, dtype=torch.float64)
            h = torch.randn(M, M, generator=gen, dtype=torch.float64)
        elif M <= 64:
            g_full = synthetic.eri_spatial(M)
        else:
            g_full = None
        # ---- reference values: the oracle where it is cheap, else the unsharded engine ----
        if g_full is not None:
            hn, gn, Dn, Gn, Un = (t.numpy() for t in (h, g_full, D, G, U))
            E_ref = onp.rotated_energy_spatial(Un, Dn, Gn, hn, gn)
            g_ref = onp.rotated_energy_grad_spatial(Un, Dn, Gn, hn, gn)
            o_ref = onp.optimal_rotation(lambda X: onp.rotated_energy_spatial(X, Dn, Gn, hn, gn),
                                         lambda X: onp.rotated_energy_grad_spatial(X, Dn, Gn, hn, gn),
                                         Un, 0.02, 1e-9, 40)
            against = "oracle"
        else:
            full = esoo_b200.OrbitalEngine(M, N, device=dev)
            full.set_integrals(h, synthetic.eri_spatial(M, device=dev))
            full.set_rdms(D, G)
            E_t, g_t = full.energy_grad(U)
            E_ref, g_ref = float(E_t), g_t.cpu().numpy()
            o_ref = full.optimize(U.numpy(), 0.02, 1e-9, 40)
            full.close()
            against = "unsharded engine"
        for ar in ("nccl", "fused"):
            eng = esoo_b200.OrbitalEngine(M, N, device=dev, t0=t0, mloc=mloc)
            if mode == "packed":
                eng.set_integrals_packed(h, synthetic.eri_spatial_pair_packed(M, t0, mloc, device=dev))
            elif mode == "generic":
                eng.set_integrals(h, g_full[t0:t0 + mloc],
                                  g_pair_transposed=g_full.permute(2, 3, 0, 1)[t0:t0 + mloc].contiguous())
            else:
                eng.set_integrals(h, synthetic.eri_spatial_shard(M, t0, mloc, device=dev),
                                  assume_v4_symmetric=True)
                eng.set_pair_symmetry(mode == "pair")
            eng.set_rdms(D, G)
            esoo_b200.attach_nccl(eng)
            if ar == "fused":
                esoo_b200.attach_peer_memory(eng)
            for _ in range(3):               # several evaluations: exercises both flag parities
                E, g = eng.energy_grad(U)
            dE = abs(float(E) - E_ref) / max(1.0, abs(E_ref))
            dg = float(np.linalg.norm(g.cpu().numpy() - g_ref) / np.linalg.norm(g_ref))
            # identical bits on every rank in fused mode (fixed summation order)
            vec = torch.cat([g.reshape(-1), E.reshape(1)])
            allv = [torch.zeros_like(vec) for _ in range(world)]
            dist.all_gather(allv, vec)
            same = all(torch.equal(allv[0], v) for v in allv)
            o = eng.optimize(U.numpy(), 0.02, 1e-9, 40)
            dopt = abs(o["energy"] - o_ref["energy"])
            dU = float(np.max(np.abs(o["U"] - o_ref["U"])))
            good = dE <= E_TOL and dg <= G_RTOL and o["n_iter"] == o_ref["n_iter"] and \
                dopt <= EFINAL_TOL and dU <= 1e-6 and (same or ar == "nccl")
            if ar == "fused":
                good = good and eng.peer_status() == 1
            log(rank, f"M={M} N={N} {mode} {ar} vs {against}: dE={dE:.2e} dgrad={dg:.2e} "
                      f"bit_identical_across_ranks={same} opt n_iter={o['n_iter']}/{o_ref['n_iter']} "
                      f"dE_opt={dopt:.2e} dU_opt={dU:.2e} -> {'ok' if good else 'FAIL'}")
            ok = ok and good
            eng.close()
    return ok


class _Solver:
    wavefunction_real = True

    def __init__(self, weights=None):
        if weights is not None:
            self.weight_vector = list(weights)

    def compute_rotated_energy(self, *a, **k):
        raise AssertionError("the CUDA optimiser must not call the Python objective")

    def compute_rotated_weighted_energy_sum(self, *a, **k):
        raise AssertionError("the CUDA optimiser must not call the Python objective")


def class_cases(rank, world, dev):
    """The drop-in class under torch.distributed (SURVEY 8e behind the reference's API)."""
    from conftest import golden_inputs, load_golden
    ok = True
    # ---- complete outer loop, spin-orbital tensors on the host and on the device ----
    gold = load_golden("outer_H4_631G_ground")
    hs, gs = torch.from_numpy(gold["h_spin"]), torch.from_numpy(gold["g_spin"])
    N = int(gold["N"])
    if world <= int(gold["M"]):
        for on_host in (True, False):
            esoo_b200.clear_engine_cache()
            opt = esoo_b200.PartialUnitaryProjectionOptimizer(
                float(gold["bb0"]), float(gold["tol"]), int(gold["maxiter"]), device=str(dev),
                inputs_on_host=on_host)
            res = harness.run_outer_loop(opt, hs, gs, 2 * N, int(gold["n_alpha"]),
                                         int(gold["n_beta"]), maxiter=int(gold["outer_maxiter"]),
                                         stopping_tolerance=float(gold["outer_tol"]),
                                         engine_for_transform=opt)
            E = np.array(res["energies"])
            good = E.shape == gold["energies"].shape and \
                float(np.max(np.abs(E - gold["energies"]))) <= EFINAL_TOL
            dev_max = float(np.max(np.abs(E - gold["energies"]))) if E.shape == gold["energies"].shape else -1
            from esoo_b200 import optimizer as om
            entry = next(iter(om._ENGINE_CACHE.values()))
            sharded = entry.engine.mloc < entry.engine.M and entry.engine.world == world
            good = good and sharded
            log(rank, f"class outer loop H4/6-31G inputs_on_host={on_host}: {len(E)} outer iterations, "
                      f"max |dE| vs live-reference fixture = {dev_max:.2e}, engine rows "
                      f"[{entry.engine.t0},{entry.engine.t0 + entry.engine.mloc}) of {entry.engine.M} "
                      f"-> {'ok' if good else 'FAIL'}")
            ok = ok and good
    # ---- inner-loop trajectory fixtures through the sharded class ----
    for name in ("abba_M12_N4", "weighted_M6_N2_k3", "abab_M5_N2"):
        gold = load_golden(name)
        if world > int(gold["M"]):
            continue
        hs, gs, Ds, Gs, U0 = golden_inputs(gold)
        esoo_b200.clear_engine_cache()
        if int(gold["n_states"]) == 1:
            fun, d_arg, g_arg = _Solver().compute_rotated_energy, Ds[0], Gs[0]
        else:
            fun, d_arg, g_arg = _Solver(gold["weights"]).compute_rotated_weighted_energy_sum, Ds, Gs
        calls = []
        opt = esoo_b200.PartialUnitaryProjectionOptimizer(
            float(gold["opt_bb0"]), float(gold["opt_tol"]), int(gold["opt_maxiter"]),
            callback=lambda it, e: calls.append(it), device=str(dev))
        U, E = opt.compute_optimal_rotation(fun=fun, initial_partial_unitary=U0.clone(),
                                            oneRDM=d_arg, twoRDM=g_arg, one_body_integrals=hs,
                                            two_body_integrals=gs)
        dE = abs(float(E) - float(gold["opt_E"]))
        dU = float(np.max(np.abs(U.numpy() - gold["opt_U"])))
        good = dE <= EFINAL_TOL and calls == list(gold["opt_calls_it"]) and dU <= 1e-6
        log(rank, f"class compute_optimal_rotation {name}: {len(calls)} callbacks, dE={dE:.2e} "
                  f"dU={dU:.2e} -> {'ok' if good else 'FAIL'}")
        ok = ok and good
    # ---- spin-orbital tensors WITHOUT V4 symmetry through the sharded class (generic path) ----
    M, N = 10, 3
    if world <= M:
        gen = torch.Generator().manual_seed(21)
        g_gen = 0.1 * torch.randn(M, M, M, M, generator=gen, dtype=torch.float64)
        h_gen = torch.randn(M, M, generator=gen, dtype=torch.float64)
        h_gen = 0.5 * (h_gen + h_gen.T)
        hs, gs = synthetic.spin_orbital_integrals(h_gen, g_gen, "abba")
        Dsp, Gsp = synthetic.rdms_spin(N, seed=31)
        D, G = synthetic.rdms_spatial(N, seed=31)
        U0 = synthetic.random_partial_unitary(M, N, seed=41)
        hn, gn, Dn, Gn = (t.numpy() for t in (h_gen, g_gen, D, G))
        ref = onp.optimal_rotation(lambda X: onp.rotated_energy_spatial(X, Dn, Gn, hn, gn),
                                   lambda X: onp.rotated_energy_grad_spatial(X, Dn, Gn, hn, gn),
                                   U0.numpy(), 0.01, 1e-9, 40)
        for where in ("cpu", dev):
            esoo_b200.clear_engine_cache()
            opt = esoo_b200.PartialUnitaryProjectionOptimizer(0.01, 1e-9, 40, device=str(dev))
            U, E = opt.compute_optimal_rotation(fun=_Solver().compute_rotated_energy,
                                                initial_partial_unitary=U0.clone(), oneRDM=Dsp,
                                                twoRDM=Gsp, one_body_integrals=hs.to(where),
                                                two_body_integrals=gs.to(where))
            from esoo_b200 import optimizer as om
            entry = next(iter(om._ENGINE_CACHE.values()))
            dE = abs(float(E) - ref["energy"])
            dU = float(np.max(np.abs(U.numpy() - ref["U"])))
            good = entry.engine.generic and entry.engine.mloc < M and dE <= EFINAL_TOL and \
                opt.last_result["n_iter"] == ref["n_iter"] and dU <= 1e-6
            log(rank, f"class generic (no V4 symmetry) M={M} N={N}, tensors on {where}: n_iter="
                      f"{opt.last_result['n_iter']}/{ref['n_iter']} dE={dE:.2e} dU={dU:.2e} -> "
                      f"{'ok' if good else 'FAIL'}")
            ok = ok and good
    # ---- SpatialIntegrals shards (the format of configs 4 and 5), pair-packed ----
    M, N = 48, 8
    t0, mloc = esoo_b200.shard_range(M, rank, world)
    h = synthetic.h_spatial(M)
    g_full = synthetic.eri_spatial(M)
    Dsp, Gsp = synthetic.rdms_spin(N)
    D, G = synthetic.rdms_spatial(N)
    U0 = synthetic.random_partial_unitary(M, N)
    hn, gn, Dn, Gn = (t.numpy() for t in (h, g_full, D, G))
    ref = onp.optimal_rotation(lambda X: onp.rotated_energy_spatial(X, Dn, Gn, hn, gn),
                               lambda X: onp.rotated_energy_grad_spatial(X, Dn, Gn, hn, gn),
                               U0.numpy(), 0.02, 1e-9, 60)
    esoo_b200.clear_engine_cache()
    sp = esoo_b200.SpatialIntegrals(synthetic.eri_spatial_pair_packed(M, t0, mloc, device=dev), M,
                                    t0=t0, mloc=mloc, packed=True)
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(0.02, 1e-9, 60, device=str(dev))
    U, E = opt.compute_optimal_rotation(fun=_Solver().compute_rotated_energy,
                                        initial_partial_unitary=U0.clone(), oneRDM=Dsp, twoRDM=Gsp,
                                        one_body_integrals=h, two_body_integrals=sp)
    dE = abs(float(E) - ref["energy"])
    good = dE <= EFINAL_TOL and opt.last_result["n_iter"] == ref["n_iter"]
    log(rank, f"class SpatialIntegrals packed shard M={M} N={N} vs oracle: n_iter="
              f"{opt.last_result['n_iter']}/{ref['n_iter']} dE={dE:.2e} -> {'ok' if good else 'FAIL'}")
    ok = ok and good
    esoo_b200.clear_engine_cache()
    return ok


def config5_through_the_class(rank, world, dev):
    """BASELINE.json configs[4] (M=400, N=24, 204.8 GB dense: needs the GPUs of the box) driven
    through PartialUnitaryProjectionOptimizer.compute_optimal_rotation with SpatialIntegrals shards.
    No CPU answer exists at this size (the oracle checks thin M=400 shards in tests/test_gpu_parity);
    here: the ranks agree bit for bit, the callbacks arrive in order, the energy goes down."""
    import time
    M, N, maxiter = 400, 24, 12
    t0, mloc = esoo_b200.shard_range(M, rank, world)
    h = synthetic.h_spatial(M)
    Dsp, Gsp = synthetic.rdms_spin(N)
    U0 = synthetic.random_partial_unitary(M, N)
    esoo_b200.clear_engine_cache()
    g = synthetic.eri_spatial_pair_packed(M, t0, mloc, device=dev)
    sp = esoo_b200.SpatialIntegrals(g, M, t0=t0, mloc=mloc, packed=True)
    calls = []
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(1e-4, 0.0, maxiter, device=str(dev),
                                                      callback=lambda it, e: calls.append((it, e)))
    torch.cuda.synchronize()
    dist.barrier()
    t_a = time.perf_counter()
    U, E = opt.compute_optimal_rotation(fun=_Solver().compute_rotated_energy,
                                        initial_partial_unitary=U0.clone(), oneRDM=Dsp, twoRDM=Gsp,
                                        one_body_integrals=h, two_body_integrals=sp)
    t_first = time.perf_counter() - t_a
    first_calls = list(calls)
    calls.clear()
    opt.BBstepsize = 1e-4       # like the reference, the first call left its last BB step behind
    t_a = time.perf_counter()
    U2, E2 = opt.compute_optimal_rotation(fun=_Solver().compute_rotated_energy,
                                          initial_partial_unitary=U0.clone(), oneRDM=Dsp,
                                          twoRDM=Gsp, one_body_integrals=h, two_body_integrals=sp)
    t_second = time.perf_counter() - t_a
    vec = torch.cat([U.reshape(-1), E.reshape(1)]).to(dev)
    allv = [torch.zeros_like(vec) for _ in range(world)]
    dist.all_gather(allv, vec)
    same = all(torch.equal(allv[0], v) for v in allv)
    Es = [e for _, e in calls]
    n_it = opt.last_result["n_iter"]
    orth_err = float(np.max(np.abs(U.numpy().T @ U.numpy() - np.eye(N))))
    checks = {"identical_on_all_ranks": same,
              "callback_iterations": [c[0] for c in calls] == list(range(maxiter + 1)),
              "repeatable": calls == first_calls and bool(torch.equal(U, U2)),
              "n_iter": n_it == maxiter + 1, "energy_decreases": min(Es) < Es[0],
              "orthonormal": orth_err <= 1e-11}
    good = all(checks.values())
    log(rank, f"class config 5 (M={M}, N={N}, {world} GPUs, {g.numel() * 8 / 1e9:.1f} GB/GPU pair-packed): "
              f"{n_it} iterations, E {Es[0]:.6f} -> {Es[-1]:.6f}, identical on all ranks={same}, "
              f"first call {t_first:.3f} s, second call (engine cached) {t_second:.3f} s = "
              f"{n_it / t_second:.1f} iterations/s, |U^T U - I| = {orth_err:.1e}, checks {checks} -> "
              f"{'ok' if good else 'FAIL'}")
    esoo_b200.clear_engine_cache()
    return good


def peer_timeout_case(rank, world, dev):
    """A rank that waits in vain inside the fused all-reduce must fail loudly: rank 0
    evaluates once more than its peers with a 0.3 s time-out; the call has to raise and the result
    on the device has to be NaN, never a plausible-looking partial sum."""
    import time
    from esoo_b200._lib import OOError
    M, N = 32, 4
    t0, mloc = esoo_b200.shard_range(M, rank, world)
    eng = esoo_b200.OrbitalEngine(M, N, device=dev, t0=t0, mloc=mloc)
    eng.set_integrals(synthetic.h_spatial(M), synthetic.eri_spatial_shard(M, t0, mloc, device=dev),
                      assume_v4_symmetric=True)
    eng.set_rdms(*synthetic.rdms_spatial(N))
    esoo_b200.attach_nccl(eng)
    esoo_b200.attach_peer_memory(eng)
    U = synthetic.random_partial_unitary(M, N)
    E, _ = eng.energy_grad(U)                       # everybody takes part: fine
    good = bool(np.isfinite(float(E))) and eng.peer_status() == 1
    dist.barrier()
    raised, poisoned = True, True
    if rank == 0:
        eng.set_peer_timeout(0.3)
        raised = False
        try:
            eng.energy_grad(U)                      # the peers never come
        except OOError as exc:
            raised = "timed out" in str(exc)
        poisoned = bool(torch.isnan(eng._out).all())
    else:
        time.sleep(1.5)
    dist.barrier()
    flag = torch.tensor([1 if (good and raised and poisoned) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    log(rank, f"fused all-reduce time-out: call raised={raised}, device result NaN={poisoned} -> "
              f"{'ok' if int(flag.item()) == 1 else 'FAIL'}")
    eng.close()
    return int(flag.item()) == 1


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = engine_cases(rank, world, dev)
    if os.environ.get("OO_MG_SKIP_CLASS") is None:
        ok = class_cases(rank, world, dev) and ok
    if os.environ.get("OO_MG_CONFIG5") is not None:
        ok = config5_through_the_class(rank, world, dev) and ok
    ok = peer_timeout_case(rank, world, dev) and ok          # last: it leaves the peers out of step
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU PASS" if int(flag.item()) == 1 else "MULTIGPU FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
