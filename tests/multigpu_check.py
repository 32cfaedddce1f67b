"""Multi-GPU parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/multigpu_check.py

Every rank holds a shard of g (first index), evaluates, all-reduces (NCCL, then the all-reduce
fused into the tail kernel over NVLink peer memory) and compares with an unsharded engine on the
same GPU; then runs the device-resident optimiser sharded and unsharded.  Prints MULTIGPU PASS.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import esoo_b200  # noqa: E402
from esoo_b200 import synthetic  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    # (M, N, slab mode): True = pair-symmetric, False = dense, "packed" = pair-packed storage
    cases = [(64, 16, True), (64, 16, False), (50, 5, True), (272, 8, True), (64, 16, "packed")]
    cases = cases[:int(os.environ.get("OO_MG_CASES", len(cases)))]
    for (M, N, pair) in cases:
        h = synthetic.h_spatial(M)
        D, G = synthetic.rdms_spatial(N)
        U = synthetic.random_partial_unitary(M, N)
        t0, mloc = esoo_b200.shard_range(M, rank, world)
        gsh = synthetic.eri_spatial_shard(M, t0, mloc, device=dev)
        full = esoo_b200.OrbitalEngine(M, N, device=dev)
        full.set_integrals(h, synthetic.eri_spatial(M, device=dev))
        full.set_rdms(D, G)
        full.set_pair_symmetry(bool(pair))
        E_ref, g_ref = full.energy_grad(U)
        o_ref = full.optimize(U.numpy(), 0.02, 1e-9, 40)
        for mode in ("nccl", "fused"):
            eng = esoo_b200.OrbitalEngine(M, N, device=dev, t0=t0, mloc=mloc)
            if pair == "packed":
                eng.set_integrals_packed(h, synthetic.eri_spatial_pair_packed(M, t0, mloc,
                                                                              device=dev))
            else:
                eng.set_integrals(h, gsh, assume_v4_symmetric=True)
                eng.set_pair_symmetry(pair)
            eng.set_rdms(D, G)
            esoo_b200.attach_nccl(eng)
            if mode == "fused":
                esoo_b200.attach_peer_memory(eng)
            for rep in range(3):            # several evaluations: exercises both flag parities
                E, g = eng.energy_grad(U)
            if os.environ.get("OO_MG_VERBOSE"):
                print(f"[rank {rank}] {mode}: evaluations done, peer_status="
                      f"{eng.peer_status() if mode == 'fused' else '-'}", flush=True)
            dE = abs(float(E) - float(E_ref))
            dg = float((g - g_ref).norm() / g_ref.norm())
            # identical bits on every rank in fused mode (fixed summation order)
            vec = torch.cat([g.reshape(-1), E.reshape(1)])
            allv = [torch.zeros_like(vec) for _ in range(world)]
            dist.all_gather(allv, vec)
            same = all(torch.equal(allv[0], v) for v in allv)
            o = eng.optimize(U.numpy(), 0.02, 1e-9, 40)
            if os.environ.get("OO_MG_VERBOSE"):
                print(f"[rank {rank}] {mode}: optimize done n_iter={o['n_iter']}", flush=True)
            dopt = abs(o["energy"] - o_ref["energy"])
            good = dE <= 1e-10 * max(1.0, abs(float(E_ref))) and dg <= 1e-9 and \
                o["n_iter"] == o_ref["n_iter"] and dopt <= 1e-8 and (same or mode == "nccl")
            if mode == "fused":
                good = good and eng.peer_status() == 1
            if rank == 0:
                print(f"M={M} N={N} pair_sym={pair} {mode}: dE={dE:.2e} dgrad={dg:.2e} "
                      f"bit_identical_across_ranks={same} opt n_iter={o['n_iter']}/{o_ref['n_iter']} "
                      f"dE_opt={dopt:.2e} -> {'ok' if good else 'FAIL'}", flush=True)
            ok = ok and good
            eng.close()
        full.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU PASS" if int(flag.item()) == 1 else "MULTIGPU FAIL", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
