"""Seeded synthetic workloads shared by the parity tests and bench.py."""
import numpy as np

# X, U, ref, active rows, dU*[0]
KNOWN_CASES = [
    dict(name="A", X=[.01, 0, .02, 0], U=0.0, ref=0.0, active=[], du0=1.998944554495782e-05,
         obj=-1.1182268817754822e-08, extra={14: -5.4550407656846813e-07}),
    dict(name="B", X=[0, 0, .05, 0], U=0.0, ref=0.0, active=list(range(15, 25)), du0=-0.4,
         obj=0.8779975093518252, extra={10: 0.01241466386295066}),
    dict(name="C", X=[.02, -.1, .03, .2], U=0.5, ref=0.25, active=list(range(15, 25)), du0=-0.38,
         obj=0.7761542266326437, extra={10: 0.00851069244123295}),
    dict(name="D", X=[.1, .5, .08, -.2], U=-1.0, ref=-0.3, active=[], du0=0.004931199246415284,
         obj=-0.0029653025550182397, extra={14: 0.00617697308408294}),
]


def c2_batch(B=4096, seed=0):
    """Config 2 inputs: X ~ N(0, diag(.05,.2,.05,.3)^2), U ~ N(0, 2^2), ref ~ N(0, .25^2)."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((B, 4)) * np.array([0.05, 0.2, 0.05, 0.3])
    U = rng.standard_normal(B) * 2.0
    ref = rng.standard_normal(B) * 0.25
    return X, U, ref


def random_qp(n, m, seed, density=1.0, active_frac=0.3):
    """Strictly convex random QP with finite two-sided bounds, some of them active at the optimum."""
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    P = M @ M.T / n + 0.5 * np.eye(n)
    A = rng.standard_normal((m, n))
    if density < 1.0:
        A *= rng.random((m, n)) < density
    q = rng.standard_normal(n)
    xs = np.linalg.solve(P, -q)
    Ax = A @ xs
    w = 0.2 + rng.random(m)
    shift = np.where(rng.random(m) < active_frac, 1.0, -1.0) * 0.3 * rng.random(m)
    u = Ax + w * 0.5 - shift
    l = u - w - 1.0
    return P, q, A, l, u


def c4_plants(B, cfg, seed=0):
    """Config 4 inputs: Ad_i = Ad + 0.01 |Ad| o N(0,1), Bd_i = -(Ad_i - I) e1, rejected if rho(Ad_i) >= 1."""
    rng = np.random.default_rng(seed)
    Ad, nx = cfg["Ad"], cfg["Ad"].shape[0]
    out_A, out_B = np.empty((B, nx, nx)), np.empty((B, nx))
    k = 0
    while k < B:
        Ai = Ad + 0.01 * np.abs(Ad) * rng.standard_normal((nx, nx))
        if np.abs(np.linalg.eigvals(Ai)).max() >= 1.0:
            continue
        out_A[k] = Ai
        out_B[k] = -(Ai - np.eye(nx))[:, 0]
        k += 1
    return out_A, out_B


def c3_batch(B, seed=0):
    """Config 3 inputs: x0 ~ U(-0.5, 0.5) on the positions, 0 elsewhere; reference = hover at z in U(0, 1)."""
    rng = np.random.default_rng(seed)
    x0 = np.zeros((B, 12))
    x0[:, :3] = rng.uniform(-0.5, 0.5, (B, 3))
    xr = np.zeros((B, 12))
    xr[:, 2] = rng.uniform(0.0, 1.0, B)
    return x0, xr
