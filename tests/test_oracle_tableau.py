"""The oracle's Tsit5 constants satisfy the Runge-Kutta order conditions, so a mis-recalled digit cannot hide.
(OrdinaryDiffEqTsit5 1.1.0 is not vendored in the reference)"""
import numpy as np

from conftest import lv_chain
from oracle import Oracle


def _orc():
    return Oracle(lv_chain().desc())


def test_tableau_row_sums_and_order5():
    c6, a, bt = _orc().tableau()
    c = np.concatenate([[0.0], c6])                      # c for stages 1..7
    A = a                                                # row s = stage s+1 (row 0 = zeros)
    assert np.allclose(A.sum(axis=1), c, atol=1e-14)
    b = A[6].copy()                                      # FSAL: a7j = b_j, b7 = 0
    one = np.ones(7)
    Ac = A @ c
    # order 1..5 conditions (17 trees)
    conds = [
        (b @ one, 1), (b @ c, 1 / 2), (b @ c**2, 1 / 3), (b @ Ac, 1 / 6), (b @ c**3, 1 / 4),
        (b @ (c * Ac), 1 / 8), (b @ (A @ c**2), 1 / 12), (b @ (A @ Ac), 1 / 24), (b @ c**4, 1 / 5),
        (b @ (c**2 * Ac), 1 / 10), (b @ (Ac * Ac), 1 / 20), (b @ (c * (A @ c**2)), 1 / 15),
        (b @ (A @ c**3), 1 / 20), (b @ (c * (A @ Ac)), 1 / 30), (b @ (A @ (c * Ac)), 1 / 40),
        (b @ (A @ (A @ c**2)), 1 / 60), (b @ (A @ (A @ Ac)), 1 / 120),
    ]
    for got, want in conds:
        assert abs(got - want) < 2e-14, (got, want)
    # embedded error weights: sum to zero and give an order-4 method  (bhat = b - btilde)
    assert abs(bt.sum()) < 1e-15
    bh = b - bt
    for got, want in [(bh @ one, 1), (bh @ c, 1 / 2), (bh @ c**2, 1 / 3), (bh @ Ac, 1 / 6), (bh @ c**3, 1 / 4),
                      (bh @ (c * Ac), 1 / 8), (bh @ (A @ c**2), 1 / 12), (bh @ (A @ Ac), 1 / 24)]:
        assert abs(got - want) < 1e-13, (got, want)


def test_interpolant_order4_and_endpoints():
    o = _orc()
    c6, A, _ = o.tableau()
    c = np.concatenate([[0.0], c6])
    Ac = A @ c
    assert np.allclose(o.interp_weights(0.0), 0.0)
    assert np.allclose(o.interp_weights(1.0), A[6], atol=1e-13)      # b_i(1) = b_i
    for th in np.linspace(0.05, 0.95, 10):
        b = o.interp_weights(th)
        for got, want in [(b.sum(), th), (b @ c, th**2 / 2), (b @ c**2, th**3 / 3), (b @ Ac, th**3 / 6),
                          (b @ c**3, th**4 / 4), (b @ (c * Ac), th**4 / 8), (b @ (A @ c**2), th**4 / 12),
                          (b @ (A @ Ac), th**4 / 24)]:
            assert abs(got - want) < 1e-13, (th, got, want)


def test_fastpower_matches_pow_to_its_published_accuracy():
    o = _orc()
    xs = np.exp(np.random.default_rng(0).uniform(np.log(1e-8), np.log(1e3), 2000))
    for y in (7 / 50, 2 / 25):
        got = np.array([o.fastpower(x, y) for x in xs])
        assert np.max(np.abs(got / xs**y - 1)) < 2e-4      # Goldberg rational log2: ~1e-4 relative
    assert o.fastpower(0.0, 0.14) == 0.0
    assert abs(o.fastpower(1.0, 0.14) - 1.0) < 1e-6
