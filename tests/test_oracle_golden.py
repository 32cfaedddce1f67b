"""The numpy oracle against golden vectors frozen from the live reference (CPU only)."""
import numpy as np
import pytest

from conftest import golden_inputs, golden_names, load_golden
from oracle import oracle_np as onp

SMALL = [n for n in golden_names() if "M56" not in n]


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


@pytest.mark.parametrize("name", [n for n in SMALL if "opt_E" in load_golden(n)])
def test_optimal_rotation_trajectory(name):
    """Driver loop, BB step and stopping rule: same callbacks, iteration count and final U."""
    gold = load_golden(name)
    hs, gs, Ds, Gs, U0 = golden_inputs(gold)
    from esoo_b200 import ingest
    h, g, st = ingest.reduce_integrals(hs, gs)
    D, G = ingest.reduce_rdms(Ds, Gs, st, list(gold["weights"]))
    hn, gn, Dn, Gn = h.numpy(), g.numpy(), D.numpy(), G.numpy()
    res = onp.optimal_rotation(
        lambda U: onp.rotated_energy_spatial(U, Dn, Gn, hn, gn),
        lambda U: onp.rotated_energy_grad_spatial(U, Dn, Gn, hn, gn),
        U0.numpy(), float(gold["opt_bb0"]), float(gold["opt_tol"]), int(gold["opt_maxiter"]))
    its = [c[0] for c in res["callbacks"]]
    assert its == list(gold["opt_calls_it"])
    assert abs(res["energy"] - float(gold["opt_E"])) <= 1e-8
    Es = np.array([c[1] for c in res["callbacks"]])
    assert np.max(np.abs(Es - gold["opt_calls_E"])) <= 1e-7
    assert np.max(np.abs(res["U"] - gold["opt_U"])) <= 1e-5
