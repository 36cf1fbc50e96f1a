"""Device table-driven log / exp (graniitti_b200/csrc/grb_fastmath.cuh) against glibc through the C-ABI unit-test hook.
Budget: the amplitude needs ~1e-15 absolute accuracy on ln(s_ij / s0) (it is multiplied by
alpha(t) ~ 1) and ~1e-15 relative accuracy on the exponential; both are asserted with margins below that."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_fast_log_accuracy(grb):
    rng = np.random.default_rng(3)
    x = np.concatenate([10.0 ** rng.uniform(-3, 12, 400000), rng.uniform(0.5, 2.0, 100000),
                        1.0 + (np.arange(129 * 8) / (128.0 * 8)),            # table interval edges and interiors
                        np.nextafter(1.0 + np.arange(1, 128) / 128.0, 0)])
    y = grb.test_fastmath(0, x)
    ref = np.log(x.astype(np.longdouble))
    err = np.abs((y - ref).astype(np.float64))
    bound = 1.2e-16 * (3.0 + np.abs(np.log2(x)))
    assert np.all(err <= bound), (err / bound).max()
    big = x > 4.0          # the sub-energies s_ij / s0 of the amplitudes
    assert np.max(err[big] / np.abs(np.log(x[big]))) < 2.5e-16


def test_fast_exp_accuracy(grb):
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.uniform(-60, 60, 400000), rng.uniform(-689, 689, 100000), np.linspace(-1e-3, 1e-3, 1001),
                        np.arange(-640, 641) * (np.log(2.0) / 64)])
    y = grb.test_fastmath(1, x)
    ref = np.exp(x.astype(np.longdouble))
    rel = np.abs(((y - ref) / ref).astype(np.float64))
    assert rel.max() < 3.4e-16, rel.max()        # <= 1.5 ulp


def test_fast_sincos_accuracy(grb):
    """Table-driven sincos of the Regge signature phase: absolute error <= 2e-16 where the amplitudes use it
    (|x| = pi/2 alpha' |t1 + t2| < ~50), growing only with the argument reduction (|x| ulp-level) beyond."""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-50, 50, 400000), rng.uniform(-1e-3, 1e-3, 50000), np.arange(-640, 641) * (np.pi / 64),
                        np.arange(-640, 641) * (np.pi / 64) + np.pi / 128, rng.uniform(-9e4, 9e4, 100000)])
    s, c = grb.test_fastmath(6, x), grb.test_fastmath(7, x)
    xl = x.astype(np.longdouble)
    es = np.abs((s - np.sin(xl)).astype(np.float64))
    ec = np.abs((c - np.cos(xl)).astype(np.float64))
    small = np.abs(x) <= 50
    assert max(es[small].max(), ec[small].max()) < 2.0e-16, (es[small].max(), ec[small].max())
    assert max(es.max(), ec.max()) < 1e-15
    # libm path beyond the table-reduction domain, inf / NaN
    x = np.array([1.0e5, -3.7e8, 1e22, np.inf, np.nan])
    s, c = grb.test_fastmath(6, x), grb.test_fastmath(7, x)
    with np.errstate(all="ignore"):
        assert np.allclose(s[:3], np.sin(x[:3]), rtol=0, atol=2e-16) and np.allclose(c[:3], np.cos(x[:3]), rtol=0, atol=2e-16)
    assert np.all(np.isnan(s[3:])) and np.all(np.isnan(c[3:]))


def test_fast_math_out_of_domain_takes_libm_path(grb):
    x = np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310, 2.2250738585072014e-308, 1.7976931348623157e308])
    y = grb.test_fastmath(0, x)
    with np.errstate(all="ignore"):
        ref = np.log(x)
    assert np.array_equal(np.isnan(y), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.allclose(y[ok], ref[ok], rtol=3e-16, atol=0) or np.array_equal(y[ok], ref[ok])
    x = np.array([-800.0, -745.2, -690.0, 690.0, 709.7, 710.0, np.inf, -np.inf, np.nan])
    y = grb.test_fastmath(1, x)
    with np.errstate(all="ignore"):
        ref = np.exp(x)
    assert np.array_equal(np.isnan(y), np.isnan(ref))
    ok = ~np.isnan(ref) & np.isfinite(ref) & (ref > 1e-300)
    assert np.allclose(y[ok], ref[ok], rtol=4e-16, atol=0)
    assert y[0] == 0.0 and np.isinf(y[5]) and np.isinf(y[6]) and y[7] == 0.0


def test_exact_sqrt_and_division_bit_identical(grb):
    """The branch-free sqrt / division sequences of the kinematic core reproduce the device's IEEE sqrt.rn / div.rn
    bit for bit (2^24 arguments per range incl. the scales of the path: GeV^2 .. s^2, plus extreme exponents)."""
    rng = np.random.default_rng(11)
    for lo, hi in ((-6, 17), (-300, 300), (-0.01, 0.01)):
        x = 10.0 ** rng.uniform(lo, hi, 1 << 24)
        assert np.array_equal(grb.test_fastmath(2, x), grb.test_fastmath(3, x))
        s = np.where(rng.random(x.size) < 0.5, -1.0, 1.0) * x
        assert np.array_equal(grb.test_fastmath(4, s), grb.test_fastmath(5, s))
    # mantissa patterns near rounding boundaries: squares and products of random 27-bit integers +- 1 ulp
    k = rng.integers(1 << 26, 1 << 27, 1 << 22).astype(np.float64)
    sq = k * k
    x = np.concatenate([sq, np.nextafter(sq, 0), np.nextafter(sq, np.inf)])
    y2, y3 = grb.test_fastmath(2, x), grb.test_fastmath(3, x)
    assert np.array_equal(y2, y3) and np.array_equal(y2[:k.size], k)
    edge = np.array([0.0, -0.0, -1.0, np.inf, np.nan, 5e-324, 1e-300, 1.7976931348623157e308, 2.0 ** -970, 2.0 ** -971])
    a, b = grb.test_fastmath(2, edge), grb.test_fastmath(3, edge)
    assert np.array_equal(a.view(np.int64), b.view(np.int64))
    a, b = grb.test_fastmath(4, edge), grb.test_fastmath(5, edge)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def test_folded_solve_pz_bit_identical(grb):
    """The loop node's SolvePz with every power-of-two factor folded into invariants (grb_dev.cuh solve_pz3_*_fast)
    reproduces the reference-order chains bit for bit (2^22 argument sets each, loop-node magnitudes at 13 TeV)."""
    rng = np.random.default_rng(21)
    x = rng.random(1 << 22)
    for a, b in ((8, 9), (10, 11)):
        ya, yb = grb.test_fastmath(a, x), grb.test_fastmath(b, x)
        assert np.isfinite(ya).mean() > 0.9
        assert np.array_equal(ya.view(np.int64), yb.view(np.int64)), np.sum(ya.view(np.int64) != yb.view(np.int64))
