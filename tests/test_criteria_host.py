"""Host-side OED helpers (mirror of lib/CorleoneOED/src/criteria.jl and oed.jl:500-534) -- no GPU needed."""
import numpy as np

import corleone_b200 as cb


def test_criteria_on_a_known_matrix():
    F = np.array([[4.0, 1.0], [99.0, 3.0]])          # Symmetric(F) reads the upper triangle: [[4, 1], [1, 3]]
    S = np.array([[4.0, 1.0], [1.0, 3.0]])
    lam = np.linalg.eigvalsh(S)
    assert np.isclose(cb.ACriterion()(F), np.trace(np.linalg.inv(S)))
    assert np.isclose(cb.DCriterion()(F), 1.0 / 11.0)
    assert np.isclose(cb.ECriterion()(F), 1.0 / lam[0])
    assert np.isclose(cb.FisherACriterion()(F), -7.0)
    assert np.isclose(cb.FisherDCriterion()(F), -11.0)
    assert np.isclose(cb.FisherECriterion()(F), -lam[0])


def test_information_gain_helpers():
    rng = np.random.default_rng(0)
    tr = rng.standard_normal((3, 5, 12))              # B = 3, 5 samples, [x(2); vec G(4); F(3); controls(3)]
    L = cb.local_information_gain(tr, 2, 2, 2)
    assert L.shape == (3, 5, 2, 2, 2)
    G = tr[1, 2, 2:6].reshape(2, 2).T                 # G[state, parameter], vec(G) column-major
    assert np.allclose(L[1, 2, 0], np.outer(G[0], G[0])) and np.allclose(L[1, 2, 1], np.outer(G[1], G[1]))
    F = np.array([[2.0, 0.3], [0.3, 1.0]])
    Gl = cb.global_information_gain(tr, F, 2, 2, 2)
    x = G[1] @ np.linalg.inv(F)
    assert np.allclose(Gl[1, 2, 1], np.outer(x, x))
    # summing the local gains over observations with unit weights gives the FIM integrand of augmentation.jl:222-227
    integrand = L[0, 0].sum(axis=0)
    G0 = tr[0, 0, 2:6].reshape(2, 2).T
    assert np.allclose(integrand, G0.T @ G0)
