import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))
                  if not os.path.basename(p).startswith(("outer_", "bb_update", "opt_decay", "fd_",
                                                         "rotated_", "bench_")))


def outer_golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "outer_*.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def golden_inputs(gold):
    """Spin-orbital inputs of a golden case as torch tensors (stored, or regenerated from seeds
    and verified against the stored checksums)."""
    import torch
    import esoo_b200
    from esoo_b200 import synthetic

    M, N, k = int(gold["M"]), int(gold["N"]), int(gold["n_states"])
    pattern = str(gold["pattern"])
    if "g_spin" in gold:
        hs, gs = torch.from_numpy(gold["h_spin"]), torch.from_numpy(gold["g_spin"])
        Ds = [torch.from_numpy(gold[f"D_spin_{n}"]) for n in range(k)]
        Gs = [torch.from_numpy(gold[f"G_spin_{n}"]) for n in range(k)]
    else:
        h = synthetic.h_spatial(M, synthetic.SEED_H)
        g = synthetic.eri_spatial(M, synthetic.SEED_ERI)
        hs, gs = synthetic.spin_orbital_integrals(h, g, pattern)
        Ds, Gs = [], []
        for n in range(k):
            D, G = synthetic.rdms_spin(N, synthetic.SEED_RDM + 17 * n)
            Ds.append(D)
            Gs.append(G)
        assert abs(float(gs.sum()) - float(gold["checksum_g"])) <= 1e-9 * max(1.0, abs(float(gold["checksum_g"])))
        assert abs(float(hs.sum()) - float(gold["checksum_h"])) <= 1e-9 * max(1.0, abs(float(gold["checksum_h"])))
    U0 = torch.from_numpy(gold["U0"])
    return hs, gs, Ds, Gs, U0


def golden_inputs_spatial(gold):
    """The same case in the spatial picture, straight from the seeds (no (2M)^4 embedding: the
    cfg-3 tensor takes 18.7 GB): (h [M,M], g [M]^4, D~, G~ spin-summed and state-weighted, U0).
    Only for cases whose inputs are regenerated; verified against the stored checksums (the
    spin-orbital tensor holds four copies of g and two of h)."""
    import torch
    from esoo_b200 import ingest, synthetic

    assert "g_spin" not in gold
    M, N, k = int(gold["M"]), int(gold["N"]), int(gold["n_states"])
    pattern = str(gold["pattern"])
    h = synthetic.h_spatial(M, synthetic.SEED_H)
    g = synthetic.eri_spatial(M, synthetic.SEED_ERI)
    assert abs(4 * float(g.sum()) - float(gold["checksum_g"])) <= 1e-9 * max(1.0, abs(float(gold["checksum_g"])))
    assert abs(2 * float(h.sum()) - float(gold["checksum_h"])) <= 1e-9 * max(1.0, abs(float(gold["checksum_h"])))
    Ds, Gs = zip(*[synthetic.rdms_spin(N, synthetic.SEED_RDM + 17 * n) for n in range(k)])
    st = ingest.SpatialIntegrals(g, M, pattern=pattern).structure
    D, G = ingest.reduce_rdms(list(Ds), list(Gs), st, list(gold["weights"]))
    return h, g, D, G, torch.from_numpy(gold["U0"])


def shard_partial_oracle(gsh, t0, U, D, G, h, pair_symmetric):
    """What ONE GPU holding rows [t0, t0+mloc) of g must put into its (gradient | energy) buffer
    (numpy oracle).  dense mode: its own rows of 4A + one-body terms.  Pair-symmetric mode: partial
    rows for every x from the slabs it streams -- slab (t,q) serves row t as is and row q
    transposed."""
    from oracle import oracle_np as onp
    from esoo_b200.distributed import pair_selected
    mloc, M = gsh.shape[0], gsh.shape[1]
    N = U.shape[1]
    Gs = 0.25 * (G + G.transpose(1, 0, 3, 2) + G.transpose(2, 3, 0, 1) + G.transpose(3, 2, 1, 0))
    rows = slice(t0, t0 + mloc)
    B1, B2 = (h @ U @ D.T)[rows], (h.T @ U @ D)[rows]
    grad = np.zeros((M, N))
    if not pair_symmetric:
        T3 = onp._transform_last3(gsh, U)
        A = np.tensordot(T3, Gs, axes=([1, 2, 3], [1, 2, 3]))
        grad[rows] = 4 * A + B1 + B2
        return grad, float(np.sum(U[rows] * (A + B1)))
    Y = np.einsum("tqrs,rk,sl->tqkl", gsh, U, U, optimize=True)       # half transform per slab
    T3 = np.zeros((M, N, N, N))
    for tl in range(mloc):
        t = t0 + tl
        for q in range(M):
            if not pair_selected(t, q):
                continue
            T3[t] += np.einsum("j,kl->jkl", U[q], Y[tl, q])
            if q != t:
                T3[q] += np.einsum("j,kl->jkl", U[t], Y[tl, q].T)
    A = np.tensordot(T3, Gs, axes=([1, 2, 3], [1, 2, 3]))
    grad = 4 * A
    grad[rows] += B1 + B2
    return grad, float(np.sum(U * A) + np.sum(U[rows] * B1))


def record_deviation(test, **values):
    """Append measured parity deviations to gpurun_out/parity_deviations.jsonl (GPU runs): the
    tolerances in the tests are set from these numbers."""
    import json
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_deviations.jsonl"), "a") as f:
            f.write(json.dumps({"test": test, **{k: float(v) for k, v in values.items()}}) + "\n")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
