"""The exactness argument of the FFMA2 Chamfer scan (pcl_chamfer.cu) rests on one inequality:
    | a(q,t) + |q|^2 - d(q,t) |  <=  2^-24 * (21 |t|^2 + 15 |q|^2)
with a = fma(tz,-2qz, fma(ty,-2qy, fma(tx,-2qx, fl|t|^2))) (what the kernel scans) and d the oracle's fp32 distance.
Checked here in numpy fp32 on random and adversarial clouds (far from the origin, mixed scales, near-duplicates), together
with the consequence the kernel uses: the true nearest neighbour's a is within E2 = 2^-18 (max|t|^2 + |q|^2) of min a."""
import numpy as np
import pytest

F = np.float32
U = 2.0 ** -24


def fma(a, b, c):  # fp32 fused multiply-add: the fp64 product of two fp32 numbers is exact
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def approx_and_exact(q, t, fused):
    tx, ty, tz = t[None, :, 0], t[None, :, 1], t[None, :, 2]
    qx, qy, qz = q[:, None, 0], q[:, None, 1], q[:, None, 2]
    w = fma(tz, tz, fma(ty, ty, (tx * tx).astype(F)))
    a = fma(tz, F(-2) * qz, fma(ty, F(-2) * qy, fma(tx, F(-2) * qx, w)))
    dx, dy, dz = (qx - tx).astype(F), (qy - ty).astype(F), (qz - tz).astype(F)
    if fused:
        d = fma(dz, dz, fma(dy, dy, (dx * dx).astype(F)))
    else:
        d = (((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F) + (dz * dz).astype(F)).astype(F)
    return a, d


def clouds():
    g = np.random.default_rng(5)
    base = lambda n: g.random((n, 3)).astype(F)
    yield "unit cube", base(300), base(700)
    for off in (3.0, 100.0, -2.5e4):
        yield f"offset {off}", (base(200) + F(off)).astype(F), (base(500) + F(off)).astype(F)
    yield "tiny", (base(200) * F(1e-3)).astype(F), (base(500) * F(1e-3)).astype(F)
    yield "large", (base(200) * F(1e6)).astype(F), (base(500) * F(1e6)).astype(F)
    t = base(600)
    yield "near duplicates", (t[:200] + (g.standard_normal((200, 3)) * 1e-6).astype(F)).astype(F), t
    yield "mixed scales", base(200), (base(500) * np.logspace(-4, 4, 500, dtype=F)[:, None]).astype(F)


@pytest.mark.parametrize("fused", [False, True])
def test_expanded_form_error_bound_and_window(fused):
    for name, q, t in clouds():
        a, d = approx_and_exact(q, t, fused)
        q2 = (q.astype(np.float64) ** 2).sum(1)[:, None]
        t2 = (t.astype(np.float64) ** 2).sum(1)[None, :]
        err = np.abs(a.astype(np.float64) + q2 - d.astype(np.float64))
        bound = U * (21 * t2 + 15 * q2)
        assert (err <= bound).all(), (name, float((err / bound).max()))
        # the window the kernel relies on: a(k*) <= min_k a + E2 for the exact nearest neighbour k*
        kstar = d.argmin(1)
        e2 = 2.0 ** -18 * (t2.max() + q2[:, 0])
        assert (a[np.arange(len(q)), kstar].astype(np.float64) <= a.min(1).astype(np.float64) + e2).all(), name
