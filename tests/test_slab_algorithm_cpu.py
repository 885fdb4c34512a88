"""The arithmetic of the slab z kernels (incompact3d_b200/csrc/x3d_slab_kernels.cuh), restated in numpy on the CPU: the periodic
compact system tri(alpha, 1, alpha) x = r solved slab by slab -- zero-carry recurrences on every slab's own rows, the carries
Yout / Z0 of the neighbouring slabs, the face corrections A(i) Yin + B(i) Zin -- against a direct solve of the whole periodic
system (what the reference's Thomas + Sherman-Morrison does, src/derive.f90:45-59).  The kernels are compared with the oracle on
the GPU (tests/test_solver_gpu.py::test_slab_z_momentum_matches_oracle, tests/multigpu_worker.py)."""
import numpy as np
import pytest


def direct(alpha, r):
    n = len(r)
    A = np.eye(n) + alpha * (np.roll(np.eye(n), 1, axis=1) + np.roll(np.eye(n), -1, axis=1))
    return np.linalg.solve(A, r)


@pytest.mark.parametrize("alpha", [1.0 / 3.0, 2.0 / 11.0])       # first / second derivative of the sixth-order schemes
@pytest.mark.parametrize("nslab,n", [(2, 256), (4, 128), (8, 64), (3, 64)])
def test_slab_solve_equals_the_periodic_solve(alpha, nslab, n):
    rng = np.random.default_rng(n + nslab)
    r = rng.standard_normal(nslab * n)
    rho = (-1.0 + np.sqrt(1.0 - 4.0 * alpha * alpha)) / (2.0 * alpha)     # alpha rho^2 + rho + alpha = 0, |rho| < 1
    c = -alpha / rho                                                      # tri(alpha, 1, alpha) = c (I - rho S-)(I - rho S+)
    assert abs(rho) ** n < 1e-17                                          # eligibility (mom_slab_eligible)
    i = np.arange(n)
    A = rho ** (i + 1) * (1.0 - rho ** (2.0 * (n - i))) / (1.0 - rho * rho)
    B = rho ** (n - i)
    z0, yout, z0first = [], [], []
    for g in range(nslab):                       # k_mom_slab: zero carries
        rs = r[g * n:(g + 1) * n]
        y = np.empty(n)
        t = 0.0
        for k in range(n):
            t = rs[k] + rho * t
            y[k] = t
        z = np.empty(n)
        t = 0.0
        for k in range(n - 1, -1, -1):
            t = y[k] + rho * t
            z[k] = t
        z0.append(z); yout.append(y[-1]); z0first.append(z[0])
    x = np.empty(nslab * n)
    for g in range(nslab):                       # carry exchange + k_zfix
        yin = yout[(g - 1) % nslab]
        zin = z0first[(g + 1) % nslab] + A[0] * yout[g]
        x[g * n:(g + 1) * n] = (z0[g] + A * yin + B * zin) / c
    ref = direct(alpha, r)
    assert np.abs(x - ref).max() < 5e-14 * np.abs(ref).max()
