"""CPU-only tests: C-ABI surface of the built library, host-side ingest / sharding logic, the
world-size-2 data path over gloo, and the "fail loudly without a GPU" contract."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import esoo_b200
    from esoo_b200 import _lib
    header = open(os.path.join(ROOT, "include", "oo_b200.h")).read()
    declared = set(re.findall(r"\b(oo_[a-z0-9_]+)\s*\(", header))
    declared -= {"oo_ctx", "oo_status"}
    assert declared, "no declarations found in include/oo_b200.h"
    assert os.path.isfile(_lib.LIB_PATH), "liboo_b200.so has not been built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"symbols declared in the header but not exported: {missing}"
    assert declared == set(_lib.SYMBOLS)
    assert b"sm_100a" in _lib.load().oo_version()


def test_no_cpu_fallback():
    import esoo_b200
    with pytest.raises(ValueError):
        esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 10, device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            esoo_b200.OrbitalEngine(8, 2, device="cuda:0")
        # the C layer reports the missing device instead of computing anything
        lib = esoo_b200._lib.load()
        ctx = ctypes.c_void_p()
        rc = lib.oo_create(0, 8, 2, 0, 8, ctypes.byref(ctx))
        assert rc != 0 and lib.oo_last_error()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "electronic-structure-orbital-optimization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_constructor_mirrors_reference_signature():
    import inspect
    import esoo_b200
    sig = inspect.signature(esoo_b200.PartialUnitaryProjectionOptimizer.__init__)
    # the reference's parameters in the reference's order; extensions only after them
    assert list(sig.parameters)[1:8] == ["initial_BBstepsize", "stopping_tolerance", "maxiter",
                                         "callback", "decay_factor", "gradient_method", "device"]
    assert list(sig.parameters)[8:] == ["inputs_on_host", "distributed", "cache_check"]
    assert sig.parameters["inputs_on_host"].default is False
    assert sig.parameters["distributed"].default is None
    assert sig.parameters["cache_check"].default == "full"
    oh = esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 7, device="cuda:1",
                                                     inputs_on_host=True)
    assert oh.device == "cpu" and oh.compute_device == "cuda:1"
    o = esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 7, device="cuda:1")
    assert (o.BBstepsize, o.stopping_tolerance, o.maxiter, o.decay_factor, o.device,
            o.gradient_method, o.callback) == (0.1, 1e-6, 7, 0.8, "cuda:1", "autograd", None)
    import copy
    c = copy.deepcopy(o)            # base_opt_orb_solver.py:75 deep-copies the optimiser
    c.BBstepsize = 0.5
    assert o.BBstepsize == 0.1
    for m in ("orth", "compute_rotated_energy_automatic_gradient", "compute_rotated_energy_gradient",
              "compute_updated_partial_unitary", "compute_optimal_rotation"):
        assert callable(getattr(o, m))
    ref_sig = ["fun", "initial_partial_unitary", "oneRDM", "twoRDM", "one_body_integrals",
               "two_body_integrals"]
    assert list(inspect.signature(o.compute_optimal_rotation).parameters) == ref_sig


def test_signature_matches_live_reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present on this machine")
    import inspect
    import esoo_b200
    Ref, _ = ref_loader.load_reference()
    ours = esoo_b200.PartialUnitaryProjectionOptimizer
    for name in ("__init__", "orth", "compute_rotated_energy_automatic_gradient",
                 "compute_rotated_energy_gradient", "compute_updated_partial_unitary",
                 "compute_optimal_rotation"):
        mine = list(inspect.signature(getattr(ours, name)).parameters)
        theirs = list(inspect.signature(getattr(Ref, name)).parameters)
        if name == "__init__":                      # extensions may follow the reference's list
            assert mine[:len(theirs)] == theirs and \
                mine[len(theirs):] == ["inputs_on_host", "distributed", "cache_check"]
        else:
            assert mine == theirs, name


def test_fun_identification():
    from functools import partial
    from esoo_b200.optimizer import _fun_identity

    class S:
        weight_vector = [2, 1]

        def compute_rotated_energy(self):
            pass

        def compute_rotated_weighted_energy_sum(self):
            pass

        def other(self):
            pass

    s = S()
    assert _fun_identity(s.compute_rotated_energy) == ("compute_rotated_energy", s)
    assert _fun_identity(partial(s.compute_rotated_weighted_energy_sum))[0] == \
        "compute_rotated_weighted_energy_sum"
    with pytest.raises(TypeError):
        _fun_identity(s.other)
    with pytest.raises(TypeError):
        _fun_identity(lambda: 0)


@pytest.mark.parametrize("pattern", ["abba", "abab"])
def test_ingest_detects_block_pattern(pattern):
    from esoo_b200 import ingest, synthetic
    M, N = 5, 2
    h, g = synthetic.h_spatial(M), synthetic.eri_spatial(M)
    hs, gs = synthetic.spin_orbital_integrals(h, g, pattern)
    h2, g2, st = ingest.reduce_integrals(hs, gs)
    assert torch.equal(h2, h) and torch.equal(g2, g) and len(st.blocks) == 4
    D, G = synthetic.rdms_spin(N)
    Dsp, Gsp = ingest.reduce_rdms(D, G, st)
    Dref, Gref = synthetic.rdms_spatial(N, pattern=pattern)
    assert torch.allclose(Dsp, Dref, atol=0, rtol=0) and torch.allclose(Gsp, Gref, atol=1e-15)
    # weights: linear combination of states
    D2, G2 = synthetic.rdms_spin(N, seed=5)
    Dw, Gw = ingest.reduce_rdms([D, D2], [G, G2], st, [2.0, 1.0])
    Da, Ga = ingest.reduce_rdms(D2, G2, st)
    assert torch.allclose(Dw, 2 * Dsp + Da, atol=1e-15) and torch.allclose(Gw, 2 * Gsp + Ga, atol=1e-15)


def test_ingest_rejects_out_of_contract_inputs():
    from esoo_b200 import ingest, synthetic
    M, N = 4, 2
    h, g = synthetic.h_spatial(M), synthetic.eri_spatial(M)
    hs, gs = synthetic.spin_orbital_integrals(h, g)
    bad = gs.clone()
    bad[:M, :M, :M, :M] = 2 * g                       # unrestricted: alpha-alpha block differs
    with pytest.raises(NotImplementedError):
        ingest.reduce_integrals(hs, bad)
    hb = hs.clone()
    hb[0, M] = 1.0                                    # alpha-beta coupling
    with pytest.raises(NotImplementedError):
        ingest.reduce_integrals(hb, gs)
    with pytest.raises(TypeError):
        ingest.reduce_integrals(hs.float(), gs.float())
    _, _, st = ingest.reduce_integrals(hs, gs)
    D, G = synthetic.rdms_spin(N)
    with pytest.raises(NotImplementedError):
        ingest.reduce_rdms(D.to(torch.complex128), G.to(torch.complex128), st)
    with pytest.raises(ValueError):
        ingest.reduce_rdms([D, D], [G, G], st, [1.0])


def test_shard_range_partitions():
    from esoo_b200 import shard_range
    for M, W in [(256, 8), (400, 8), (10, 4), (7, 7), (28, 1), (110, 3)]:
        rows = []
        for r in range(W):
            t0, n = shard_range(M, r, W)
            assert n >= 1
            rows += list(range(t0, t0 + n))
        assert rows == list(range(M))
    with pytest.raises(ValueError):
        shard_range(3, 0, 4)


def test_synthetic_eri_is_v4_symmetric_and_shards_agree():
    from esoo_b200 import synthetic
    M = 9
    g = synthetic.eri_spatial(M)
    for perm in [(1, 0, 3, 2), (2, 3, 0, 1), (3, 2, 1, 0)]:
        assert torch.allclose(g, g.permute(*perm), atol=1e-15)
    sh = synthetic.eri_spatial_shard(M, 3, 4)
    assert torch.equal(sh, g[3:7])
    U = synthetic.random_partial_unitary(12, 5)
    assert torch.allclose(U.T @ U, torch.eye(5, dtype=torch.float64), atol=1e-14)
    D, G = synthetic.rdms_spin(3)
    assert abs(float(torch.trace(D)) - 2.0) < 1e-12           # 1 alpha + 1 beta electron
    assert torch.allclose(G, -G.permute(1, 0, 2, 3), atol=1e-15)  # antisymmetry in (p,q)


# ------------------------------------------------------------------------------------------------
# world-size-2 data path on CPU (gloo): shard -> partial (E, grad rows) -> all-reduce == full
# ------------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import esoo_b200
    from esoo_b200 import synthetic, distributed
    from oracle import oracle_np as onp
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank,
                            world_size=world)
    M, N = 11, 3
    h, g = synthetic.h_spatial(M).numpy(), synthetic.eri_spatial(M).numpy()
    D, G = (t.numpy() for t in synthetic.rdms_spatial(N))
    U = synthetic.random_partial_unitary(M, N).numpy()
    t0, mloc = distributed.shard_range(M, rank, world)
    rows = slice(t0, t0 + mloc)
    Gs = 0.25 * (G + G.transpose(1, 0, 3, 2) + G.transpose(2, 3, 0, 1) + G.transpose(3, 2, 1, 0))
    T3 = onp._transform_last3(g[rows], U)
    A = np.tensordot(T3, Gs, axes=([1, 2, 3], [1, 2, 3]))
    B1, B2 = (h @ U @ D.T)[rows], (h.T @ U @ D)[rows]
    buf = np.zeros(M * N + 1)
    buf[:M * N].reshape(M, N)[rows] = 4 * A + B1 + B2
    buf[M * N] = np.sum(U[rows] * (A + B1))
    t = torch.from_numpy(buf)
    dist.all_reduce(t)                                   # the (M*N+1)-double sum all-reduce

    class FakeEngine:                                    # attach_nccl's bootstrap, without NCCL
        device = torch.device("cpu")
        got = None

        @staticmethod
        def nccl_unique_id():
            return bytes(range(128))

        def attach_comm(self, uid, r, w):
            FakeEngine.got = (uid, r, w)

    fe = FakeEngine()
    distributed.attach_nccl(fe)
    E_ref = onp.rotated_energy_spatial(U, D, G, h, g)
    g_ref = onp.rotated_energy_grad_spatial(U, D, G, h, g)
    ok = abs(t[M * N].item() - E_ref) < 1e-12 and \
        np.max(np.abs(t[:M * N].numpy().reshape(M, N) - g_ref)) < 1e-12 and \
        FakeEngine.got == (bytes(range(128)), rank, world)
    # ---- the host side of the sharded class path (optimizer._engine_from_spin) ----------------
    # every rank extracts only its rows of the spatial block from the reference's spin-orbital
    # tensor, the ranks agree on the non-zero spin blocks and on the V4 symmetry of their rows,
    # and the pair-symmetric partial buffers (what liboo_b200 produces per GPU) sum to the total
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import shard_partial_oracle
    from esoo_b200 import ingest
    hs, gs = synthetic.spin_orbital_integrals(torch.from_numpy(h), torch.from_numpy(g), "abba")
    h_sp, g_rows, st = ingest.reduce_integrals_rows_host(hs, gs, t0, mloc)
    ok = ok and np.array_equal(g_rows.numpy(), g[rows]) and np.array_equal(h_sp.numpy(), h)
    mask = torch.tensor([ingest.block_mask(st)], dtype=torch.int64)
    lo, hi = mask.clone(), mask.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok = ok and int(lo) == int(hi) == ingest.block_mask(ingest.reduce_integrals(hs, gs)[2])
    asym, gmax = ingest.v4_asymmetry_rows(gs, M, st.blocks[0], t0, mloc)
    ok = ok and asym <= 1e-13 * gmax
    bad = gs.clone()
    for p_, q_, r_, s_ in ((t0, 2, 3, 1), (t0, 2 + M, 3 + M, 1), (t0 + M, 2, 3, 1 + M),
                           (t0 + M, 2 + M, 3 + M, 1 + M)):
        bad[p_, q_, r_, s_] += 0.5
    ok = ok and ingest.v4_asymmetry_rows(bad, M, st.blocks[0], t0, mloc)[0] >= 0.49
    part, e_part = shard_partial_oracle(g_rows.numpy(), t0, U, D, G, h, True)
    t2 = torch.from_numpy(np.concatenate([part.reshape(-1), [e_part]]))
    dist.all_reduce(t2)
    ok = ok and abs(t2[M * N].item() - E_ref) < 1e-12 and \
        np.max(np.abs(t2[:M * N].numpy().reshape(M, N) - g_ref)) < 1e-12
    # the plan the drop-in class follows under torch.distributed (optimizer.plan_sharded_ingest)
    from esoo_b200 import optimizer as om
    plan = om.plan_sharded_ingest(hs, gs, rank, world, "cpu")
    ok = ok and plan["storage"] == "dense" and (plan["t0"], plan["mloc"]) == (t0, mloc) and \
        np.array_equal(plan["g_rows"].numpy(), g[rows])                 # M = 11 is odd: padded dense
    M2 = 8
    h8, g8 = synthetic.h_spatial(M2), synthetic.eri_spatial(M2)
    hs8, gs8 = synthetic.spin_orbital_integrals(h8, g8, "abab")
    t8, m8 = distributed.shard_range(M2, rank, world)
    plan = om.plan_sharded_ingest(hs8, gs8, rank, world, "cpu")
    ok = ok and plan["storage"] == "packed" and plan["g_pair_transposed"] is None and \
        sorted(plan["structure"].blocks) == [(0, 0, 0, 0), (0, 1, 0, 1), (1, 0, 1, 0), (1, 1, 1, 1)]
    g_ns = g8.clone()
    g_ns[M2 - 1, 0, 1, 2] += 0.25        # ONE rank sees an asymmetric row: all ranks must go generic
    hs_ns, gs_ns = synthetic.spin_orbital_integrals(h8, g_ns, "abba")
    plan = om.plan_sharded_ingest(hs_ns, gs_ns, rank, world, "cpu")
    ok = ok and plan["storage"] == "generic" and \
        torch.equal(plan["g_pair_transposed"], g_ns.permute(2, 3, 0, 1)[t8:t8 + m8].contiguous()) and \
        torch.equal(plan["g_rows"], g_ns[t8:t8 + m8])
    # pair-packed slab lists of the shards partition the full list
    mine = distributed.pair_slab_list(M, t0, mloc)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(mine)]))
    ok = ok and sum(int(c) for c in counts) == len(distributed.pair_slab_list(M))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_harness_fci_matches_energy_functional():
    """Conventions of the exact-diagonalisation harness: <psi|H(U)|psi> equals the reference's
    energy functional evaluated with the RDMs of psi (oracle restatement of base.py:554-563)."""
    from esoo_b200 import harness, synthetic
    from oracle import oracle_np as onp
    M, N = 6, 2
    h = synthetic.h_spatial(M)
    g = synthetic.eri_spatial(M, scale=0.5)
    hs, gs = synthetic.spin_orbital_integrals(h, g, "abba")
    U = synthetic.random_partial_unitary(M, N)
    sec = harness.FockSector(2 * N, 2)
    hr, gr = harness._rotated_spin_integrals(hs, gs, U)
    H = sec.hamiltonian(hr, gr)
    sub = sec.sz_subspace(1, N)
    ev, evec = np.linalg.eigh(H[np.ix_(sub, sub)])
    for n in range(2):
        psi = np.zeros(len(sec.dets))
        psi[sub] = evec[:, n]
        D, G = sec.rdms(psi)
        assert abs(np.trace(D) - 2.0) < 1e-12
        assert np.max(np.abs(G + G.transpose(1, 0, 2, 3))) < 1e-12
        E = onp.rotated_energy_spin(U.numpy(), D, G, hs.numpy(), gs.numpy())
        assert abs(E - ev[n]) < 1e-11


def test_outer_goldens_present():
    from conftest import outer_golden_names, load_golden
    names = outer_golden_names()
    assert len(names) >= 4
    for n in names:
        g = load_golden(n)
        E = g["energies"]
        assert E.shape[0] >= 2 and np.all(np.isfinite(E))
        # the orbital optimisation lowers the (state-averaged) energy at the first outer step
        w = g["weights"][:E.shape[1]]
        assert float(E[1] @ w) < float(E[0] @ w)


def test_header_is_plain_c(tmp_path):
    """include/oo_b200.h is a C ABI: it must compile as C99 with no CUDA / C++ / torch types."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "oo_b200.h"\n'
                   'int probe(void) { oo_ctx* c = 0; double e; (void)c;\n'
                   '  return oo_create(0, 8, 2, 0, 8, &c) + oo_energy_grad_host(c, 0, &e, 0); }\n')
    out = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-c", str(src), "-I",
                          os.path.join(ROOT, "include"), "-o", str(tmp_path / "abi.o")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr


def test_pair_packed_slab_list_and_generator():
    """Pair-packed storage: the host-only C entry point, its Python mirror and the slab-wise
    synthetic generator agree; the shards of the list partition the M(M+1)/2 pairs."""
    import ctypes as C
    import torch
    from esoo_b200 import _lib, distributed, synthetic
    lib = _lib.load()
    for M, world in [(10, 1), (10, 3), (17, 4), (64, 8)]:
        total = []
        for r in range(world):
            t0, mloc = distributed.shard_range(M, r, world)
            lst = distributed.pair_slab_list(M, t0, mloc)
            n = lib.oo_pair_slab_list(M, t0, mloc, None, 0)
            assert n == len(lst)
            buf = (C.c_int * (2 * n))()
            assert lib.oo_pair_slab_list(M, t0, mloc, buf, n) == n
            assert [(buf[2 * i], buf[2 * i + 1]) for i in range(n)] == lst
            assert lib.oo_pair_slab_list(M, t0, mloc, buf, n - 1) < 0      # capacity too small
            total += lst
        assert len(total) == M * (M + 1) // 2
        assert len({(min(t, q), max(t, q)) for t, q in total}) == len(total)   # one slab per pair
    assert lib.oo_pair_slab_list(8, 4, 5, None, 0) < 0
    M, t0, mloc = 12, 5, 4
    g = synthetic.eri_spatial_shard(M, t0, mloc)
    packed = synthetic.eri_spatial_pair_packed(M, t0, mloc)
    ref = torch.stack([g[t - t0, q] for t, q in distributed.pair_slab_list(M, t0, mloc)])
    assert torch.equal(packed, ref)


def test_spatial_integrals_and_checksum_host_logic():
    """Extension input format and the engine-cache fingerprint (no GPU needed)."""
    import esoo_b200
    from esoo_b200 import ingest, optimizer as om, synthetic
    M = 6
    g = synthetic.eri_spatial(M)
    sp = esoo_b200.SpatialIntegrals(g, M)
    assert (sp.t0, sp.mloc, sp.packed, sp.v4_symmetric) == (0, M, False, True)
    assert sorted(sp.structure.blocks) == [(0, 0, 0, 0), (0, 1, 1, 0), (1, 0, 0, 1), (1, 1, 1, 1)]
    assert sorted(esoo_b200.SpatialIntegrals(g, M, pattern="abab").structure.blocks) == \
        [(0, 0, 0, 0), (0, 1, 0, 1), (1, 0, 1, 0), (1, 1, 1, 1)]
    # the block mask that selects the 2-RDM blocks matches what the spin-orbital ingest detects
    hs, gs = synthetic.spin_orbital_integrals(synthetic.h_spatial(M), g, "abba")
    assert ingest.block_mask(sp.structure) == ingest.block_mask(ingest.reduce_integrals(hs, gs)[2])
    with pytest.raises(ValueError):
        esoo_b200.SpatialIntegrals(g, M, pattern="aabb")
    with pytest.raises(ValueError):
        esoo_b200.SpatialIntegrals(g, M, packed=True, v4_symmetric=False)
    with pytest.raises(ValueError):                  # a non-symmetric shard needs the transposed rows
        esoo_b200.SpatialIntegrals(g[:3], M, t0=0, mloc=3, v4_symmetric=False)
    # fingerprint: full mode sees one changed element and a permutation, sample mode may not
    a = torch.randn(7, 5, 5, 5, dtype=torch.float64)
    b = a.clone()
    b[3, 2, 1, 4] += 1e-12
    assert om.content_checksum(a) != om.content_checksum(b)
    assert om.content_checksum(a) != om.content_checksum(a.transpose(1, 2).contiguous())
    assert om.content_checksum(a) == om.content_checksum(a.clone())
    assert om.content_checksum(a, sample=True) == om.content_checksum(b, sample=True)   # blind spot
    with pytest.raises(ValueError):
        esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 10, device="cuda:0",
                                                    cache_check="maybe")
    opt = esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 10, device="cuda:0",
                                                      distributed=False)
    assert opt._ranks() == (0, 1)
    with pytest.raises(RuntimeError):
        esoo_b200.PartialUnitaryProjectionOptimizer(0.1, 1e-6, 10, device="cuda:0",
                                                    distributed=True)._ranks()


def test_expand_spin_blocks_layout():
    """Spatial h', g' -> the reference's spin-blocked Q^4 layout (base_opt_orb_solver.py:597-612)."""
    from esoo_b200 import ingest, rotated
    N = 3
    h = torch.arange(N * N, dtype=torch.float64).reshape(N, N)
    g = torch.arange(N ** 4, dtype=torch.float64).reshape(N, N, N, N) + 1.0
    st = ingest.SpinStructure(M=5, blocks=[(0, 0, 0, 0), (0, 1, 1, 0), (1, 0, 0, 1), (1, 1, 1, 1)])
    hs, gs = rotated.expand_spin_blocks(h, g, st)
    assert hs.shape == (2 * N, 2 * N) and gs.shape == (2 * N,) * 4
    assert np.array_equal(hs[:N, :N], h.numpy()) and np.array_equal(hs[N:, N:], h.numpy())
    assert not hs[:N, N:].any() and not hs[N:, :N].any()
    assert np.array_equal(gs[:N, N:, N:, :N], g.numpy()) and np.array_equal(gs[N:, N:, N:, N:], g.numpy())
    assert not gs[:N, N:, :N, N:].any()              # (a,b,a,b) is not a block of this pattern
    assert np.count_nonzero(gs) == 4 * N ** 4
