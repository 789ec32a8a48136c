"""CPU test of the lane-level model of the banded LDL^T front (tests/front_emulator.py):
the slot / lane / reload arithmetic of csrc/vbfem_front.cuh and the algebra of the merged
forward + adjoint solve (unit-vector right-hand sides, dot-product observation, one joint
back substitution) against dense NumPy solves."""
import numpy as np
import pytest

from front_emulator import Front, twisted_forward_adjoint


def rand_spd_band(rng, n, b):
    A = np.zeros((n, n))
    for i in range(n):
        for j in range(max(0, i - b), i):
            A[i, j] = A[j, i] = rng.standard_normal() * 0.3
    A += np.eye(n) * (np.abs(A).sum(1).max() + 1.0)
    return A


def test_single_front_matches_dense_ldlt():
    rng = np.random.default_rng(0)
    n, b = 100, 25
    P = b + 1
    A = rand_spd_band(rng, n, b)
    band = np.zeros((n, P))
    for hi in range(n):
        for lo in range(max(0, hi - b), hi + 1):
            band[lo, hi - lo] = A[hi, lo]
    f = rng.standard_normal(n)
    z = [f.copy()]
    F = Front(band, n, n, b, z)
    L, D, M = np.eye(n), np.zeros(n), A.copy()
    for j in range(n):
        D[j] = M[j, j]
        L[j + 1:, j] = M[j + 1:, j] / D[j]
        M[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j]) * D[j]
    for j in range(n):
        F.step(j)
        ref = np.array([1.0 / D[j]] + [L[j + t, j] if j + t < n else 0.0 for t in range(1, P)])
        assert np.abs(band[j] - ref).max() < 1e-12
    assert not F.bad
    assert np.abs(z[0] - np.linalg.solve(L, f)).max() < 1e-12   # fused forward elimination


@pytest.mark.parametrize("n,pT", [(440, 220), (440, 200), (200, 60), (131, 26), (90, 32)])
def test_twisted_forward_adjoint(n, pT):
    rng = np.random.default_rng(n + pT)
    b, P = 25, 26
    A = rand_spd_band(rng, n, b)
    f = rng.standard_normal(n)
    tip = (n - 2, n - 1)
    gy = rng.standard_normal(2)
    Q = rng.standard_normal((P, P))
    u, y, psi, bad = twisted_forward_adjoint(A, f, b, pT, tip, gy, lambda um: Q @ um)
    u0 = np.linalg.solve(A, f)
    w = np.zeros(n)
    w[tip[0]], w[tip[1]] = gy
    w[pT:pT + P] += Q @ u0[pT:pT + P]
    psi0 = np.linalg.solve(A, w)
    assert not bad
    assert np.abs(u - u0).max() < 1e-12 * np.abs(u0).max()
    assert np.abs(y - u0[list(tip)]).max() < 1e-12 * np.abs(u0).max()
    assert np.abs(psi - psi0).max() < 1e-12 * np.abs(psi0).max()
