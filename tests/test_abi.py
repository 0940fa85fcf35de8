"""The C-ABI library without a GPU: it builds, loads, exports every symbol include/fofb.h declares, and its
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "fofb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fofb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    L = C.CDLL(lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for name in names:
        assert hasattr(L, name), f"{name} is declared in include/fofb.h but not exported"


def test_params_struct_matches_header(lib):
    p = lib.default_params()
    assert p.struct_size == C.sizeof(lib.Params)
    assert (p.H0, p.Om0, p.Ode0) == (70.0, 0.3, 0.7)
    assert (p.max_velocity, p.linking_length_factor, p.virial_radius, p.richness, p.overdensity) == (2000.0, 0.2, 1.5, 12, 2.0)
    assert (p.min_window, p.min_count, p.min_virial, p.n_random, p.n_bins, p.grid_cells) == (8, 8, 12, 300, 10, 64)
    assert lib.load().fofb_abi_version() == 1


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(lib):
    with pytest.raises(lib.FofbError) as e:
        lib.Handle(0)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    from fof_b200 import fof
    from fof_b200.mock import make_catalog
    with pytest.raises(lib.FofbError):
        fof.candidate_center_search(make_catalog(100, seed=1), max_velocity=2000.0, fname="x")


def test_bad_struct_size_is_rejected(lib):
    p = lib.default_params()
    p.struct_size = 8
    out = C.c_double()
    assert lib.load().fofb_mean_separation(C.byref(p), 12.0, 1.0, C.byref(out)) == -1


def test_bootstrap_index_stream_is_numpys(lib):
    """fofb_bootstrap_indices == np.random.randint(0, n) after np.random.seed(1): the stream astropy's
    bootstrap consumes under NumpyRNGContext(1) (mass_estimator.py:178-184)."""
    L = lib.load()
    for n in (2, 3, 12, 13, 64, 65, 100, 777, 5000):
        cnt = 100 * (n - 1)
        got = np.empty(cnt, dtype=np.int32)
        assert L.fofb_bootstrap_indices(n, cnt, got.ctypes.data) == 0
        state = np.random.get_state()
        np.random.seed(1)
        want = np.concatenate([np.random.randint(low=0, high=n, size=n - 1) for _ in range(100)])
        np.random.set_state(state)
        assert np.array_equal(got, want), n


def test_vectorised_overdensity_draws_equal_the_reference_loop():
    """fof.draw_overdensity_points consumes the legacy global generator exactly like helper.py:197-202."""
    from fof_b200.fof import draw_overdensity_points
    bbox = (149.3, 150.9, 1.4, 3.0)
    np.random.seed(31)
    ra, dec = draw_overdensity_points(7, bbox, 300)
    after = np.random.random_sample()
    np.random.seed(31)
    for k in range(7):
        assert np.array_equal(ra[k], np.random.uniform(low=bbox[0], high=bbox[1], size=300))
        assert np.array_equal(dec[k], np.random.uniform(low=bbox[2], high=bbox[3], size=300))
    assert after == np.random.random_sample()
    # the unit-sample form used by fofb_pipeline_finish: low + (high - low) * u
    np.random.seed(31)
    u = np.random.random_sample((7, 2, 300))
    assert np.array_equal(bbox[0] + (bbox[1] - bbox[0]) * u[:, 0, :], ra)
    assert np.array_equal(bbox[2] + (bbox[3] - bbox[2]) * u[:, 1, :], dec)


def test_cluster_types_mirror_the_reference():
    from fof_b200.cluster import Cluster, CleanedCluster
    from fof_b200.units import Q
    center = np.arange(8.0)
    c = Cluster(center, np.zeros((13, 8)))
    assert (c.ra, c.dec, c.z, c.z_unc, c.bcg_absMag, c.gal_id, c.N) == (0.0, 1.0, 2.0, [3.0, 4.0], 5.0, 6.0, 7.0)
    assert c.richness == 13 and c.D == 0 and c.flag_poor is False
    assert np.array_equal(c.center_attributes, center)
    other = Cluster(np.array([0, 0, 0, 0, 0, -30.0, 9.0, 7.0]), np.ones((5, 8)))
    mine, theirs = c.remove_overlap(other)   # equal N, |mag| 5 < 30 -> self loses
    assert len(mine) == 0 and len(theirs) == 5
    cc = CleanedCluster(center, np.zeros((13, 8)), Q(1e14, "Msun/littleh"), Q(1e13), Q(5e5), Q(1e4), Q(0.7), -26.0)
    assert cc.properties.shape == (11,) and cc.properties[3] == 1e14
    from fof_b200.params import cosmo
    # known answer from the unmodified reference (tests/golden/cosmos_3000_default.npz holds the per-cluster values):
    # r200 = (3 M / (4 pi 200 rho_c(z)))^(1/3); at z = 2, M = 1e14 Msun/h, H0 = 70, Om0 = 0.3 -> 0.4425 Mpc/h
    r = float(cc.r200(cosmo, 200))
    rho_c = 2.77536627e11 * (0.3 * 27 + 0.7)   # h^2 Msun / Mpc^3 at z = 2
    assert abs(r - (3 * 1e14 / (4 * np.pi * 200 * rho_c)) ** (1 / 3)) < 2e-6 * r


def test_cluster_objects_pickle_like_the_reference_pipeline():
    """pipeline.py:103-104,119-120,175-176 pickles the cluster lists between stages."""
    import pickle
    from fof_b200.cluster import Cluster, CleanedCluster
    from fof_b200.units import Q
    c = Cluster(np.arange(8.0), np.ones((3, 8)))
    c.D = 2.5
    c2 = pickle.loads(pickle.dumps(c))
    assert c2.gal_id == 6.0 and c2.D == 2.5 and np.array_equal(c2.galaxies, c.galaxies)
    cc = CleanedCluster(np.arange(8.0), np.ones((3, 8)), Q(1e14, "Msun/littleh"), Q(1e13), Q(5e5), Q(1e4), Q(0.7), -26.0)
    cc2 = pickle.loads(pickle.dumps(cc))
    assert float(cc2.cluster_mass) == 1e14 and cc2.cluster_mass.unit == "Msun/littleh" and cc2.total_absmag == -26.0


def test_units_accept_plain_numbers_and_quantity_likes():
    from fof_b200.units import Q, as_float

    class FakeQuantity:  # astropy-like: has .to(unit).value
        def __init__(self, v):
            self.value = v

        def to(self, unit):
            return FakeQuantity(self.value * 1000.0) if unit == "km/s" else self

    assert as_float(2000.0) == 2000.0 and as_float(Q(1.5, "Mpc/littleh")) == 1.5
    assert as_float(FakeQuantity(2.0), "km/s") == 2000.0
    from fof_b200 import params
    assert (params.min_redshift, params.max_redshift, float(params.max_velocity), params.linking_length_factor,
            float(params.max_radius), params.richness, params.D) == (0.5, 2.5, 2000.0, 0.2, 1.5, 12, 2)


def test_r200_matches_the_unmodified_reference():
    """CleanedCluster.r200 (cluster.py:111-118) is host arithmetic: with the golden masses and centre redshifts it must
    give the radii the unmodified reference computed under the shims (delta = 200 and 500), to 1e-9."""
    from golden_util import fixtures, load
    from fof_b200.cluster import CleanedCluster
    from fof_b200.params import cosmo
    from fof_b200.units import Q
    checked = 0
    for path in fixtures():
        g = load(path)
        z_of_id = dict(zip(g["table"][:, 6].tolist(), g["table"][:, 2].tolist()))
        for cid, v, want in zip(g["cleaned_centre_ids"], g["virial"], g["r200"]):
            center = np.array([0.0, 0.0, z_of_id[cid], 0, 0, 0, cid, 0])
            cc = CleanedCluster(center, np.zeros((12, 8)), Q(v[0], "Msun/littleh"), Q(v[1]), Q(v[2]), Q(v[3]), Q(v[4]), v[5])
            np.testing.assert_allclose([float(cc.r200(cosmo, 200)), float(cc.r200(cosmo, 500))], want, rtol=1e-9)
            checked += 1
    assert checked > 50
