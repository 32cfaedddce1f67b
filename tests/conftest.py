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
                  if not os.path.basename(p).startswith(("outer_", "bb_update", "opt_decay")))


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


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
