"""Host scalar helpers of the C ABI (analysis/methods.py restated) against scipy / the oracle.
No GPU needed: these entry points are plain host code inside libfofb200.so."""
import ctypes as C

import numpy as np
import pytest

from oracle import fof_oracle as O


def _call_vec(lib, fn, params, z, *pre):
    z = np.ascontiguousarray(z, dtype=np.float64)
    out = np.empty_like(z)
    rc = fn(C.byref(params), *pre, z.size, z.ctypes.data, out.ctypes.data)
    assert rc == 0
    return out


def test_known_answers(lib):
    """SURVEY.md Appendix C (scipy quad == hyp2f1 to 5e-15)."""
    p = lib.default_params()
    L = lib.load()
    dc = _call_vec(L, L.fofb_comoving_distance, p, [0.5, 1.0, 2.0])
    np.testing.assert_allclose(dc, [1888.625395933, 3303.828805887, 5179.862074409], rtol=2e-12)
    th = np.empty(3)
    z = np.array([0.5, 1.0, 2.0])
    assert L.fofb_linear_to_angular_dist(C.byref(p), 1.5, 3, z.ctypes.data, th.ctypes.data) == 0
    np.testing.assert_allclose(th, [0.097513, 0.074324, 0.071108], atol=5e-7)


def test_distances_match_closed_form_and_quad(lib):
    """The oracle restates astropy's hyp2f1 closed form, which itself carries ~1e-14 of cancellation
    error at low z; the library's Gauss-Legendre evaluation is checked against adaptive quadrature
    to 1e-15 and against the closed form to 3e-14 over the survey (z >= 0.3)'s redshift range."""
    from scipy.integrate import quad
    p = lib.default_params()
    L = lib.load()
    z = np.concatenate([np.linspace(0.3, 3.0, 997), [5.0, 10.0, 19.9, 25.0]])
    dc = _call_vec(L, L.fofb_comoving_distance, p, z)
    ref = O.comoving_distance(z)
    assert (np.abs(dc - ref) / ref).max() < 3e-14
    zq = np.array([0.0, 1e-9, 1e-4, 0.01, 0.3, 0.5, 0.77, 1.0, 1.5, 2.0, 2.5, 7.3, 19.9, 25.0])
    dq = _call_vec(L, L.fofb_comoving_distance, p, zq)
    for zi, di in zip(zq, dq):
        q = quad(lambda x: 1.0 / float(O.efunc(x)), 0.0, zi, epsabs=0, epsrel=2e-14)[0] * O.D_H
        assert abs(di - q) <= 1e-15 * q, (zi, di, q)
    da = _call_vec(L, L.fofb_angular_diameter_distance, p, z)
    np.testing.assert_allclose(da, O.angular_diameter_distance(z), rtol=3e-14)
    for d in (0.5, 1.5):
        th = np.empty_like(z)
        assert L.fofb_linear_to_angular_dist(C.byref(p), d, z.size, z.ctypes.data, th.ctypes.data) == 0
        np.testing.assert_allclose(th, O.linear_to_angular_dist(d, z), rtol=3e-14)


def test_mean_separation_matches_quad(lib):
    p = lib.default_params()
    L = lib.load()
    for n in (12, 85, 1000, 10000):
        for z in (0.5, 0.93, 1.0, 1.7, 2.0, 2.5):
            out = C.c_double()
            assert L.fofb_mean_separation(C.byref(p), float(n), z, C.byref(out)) == 0
            ref = O.mean_separation(n, z, 2000.0, 1.7)
            assert abs(out.value - ref) / ref < 1e-13, (n, z, out.value, ref)


def test_redshift_to_velocity_bit_exact(lib):
    L = lib.load()
    rng = np.random.default_rng(3)
    z = rng.uniform(0.0, 3.0, 20000)
    for zc in (0.5, 1.2345678901234567, 2.5):
        out = np.empty_like(z)
        assert L.fofb_redshift_to_velocity(z.size, z.ctypes.data, zc, out.ctypes.data) == 0
        assert np.array_equal(out, O.redshift_to_velocity(z, zc))
