"""Cross-check of the oracle against the reference's own control flow, executed live (build container only)."""
import numpy as np
import pytest

from fof_b200.mock import make_catalog
from oracle import fof_oracle as O
from oracle import run_reference as RR

pytestmark = pytest.mark.skipif(not RR.available(), reason="/root/reference is not mounted here")


def test_live_reference_run_matches_oracle():
    import logging
    logging.disable(logging.CRITICAL)
    ref = RR.load_reference()
    df = make_catalog(2000, seed=41)
    R = RR.run_pipeline(ref, df, seed=5)
    p = O.Params()
    main_arr, cand = O.candidate_center_search(df.values, p, "fast")
    assert np.array_equal(main_arr, R["main_arr"])
    g = main_arr[(main_arr[:, 2] >= 0.5) & (main_arr[:, 2] <= 2.52)]
    c = R["candidate_centers"]
    c = c[(c[:, 2] >= 0.5) & (c[:, 2] <= 2.5)]
    np.random.seed(5)
    cl = O.FoF(g, c, p, "fast")
    assert len(cl) == len(R["clusters"])
    for x, (attrs, members, N, D) in zip(cl, R["clusters"]):
        assert np.array_equal(x.galaxies, members) and x.N == N and x.D == D
    cleaned, removed = O.interloper_removal(cl, p)
    assert removed == R["number_removed"]
    vir = O.find_masses(cleaned)
    for a, b in zip(vir, R["virial"]):
        for k in ("cluster_mass", "mass_err", "vel_disp", "vel_disp_err", "projected_radius", "total_absmag"):
            assert abs(a[k] - b[k]) <= 1e-12 * abs(b[k])


def test_shim_cosmology_known_answers():
    """SURVEY.md Appendix C."""
    ref = RR.load_reference()
    import astropy.units as u
    lad = ref.analysis_methods.linear_to_angular_dist
    for z, want in ((0.5, 0.097513), (1.0, 0.074324), (2.0, 0.071108)):
        assert abs(lad(1.5 * u.Mpc / u.littleh, z).value - want) < 5e-7
    ms = ref.analysis_methods.mean_separation
    for n, want in ((12, 6.893), (100, 3.400), (1000, 1.578), (10000, 0.733)):
        got = 0.2 * ms(n, 1.0, 0.1 * u.deg, 2000 * u.km / u.s, survey_area=1.7).value
        assert abs(got - want) < 2e-3


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_live_compare_clusters_matches_oracle(seed):
    """catalog_matching.compare_clusters, unmodified, on random samples against the oracle restatement."""
    import importlib
    import logging
    import os
    import sys
    ref = RR.load_reference()
    cwd = os.getcwd()
    os.chdir(ref.scratch)  # logging.basicConfig at catalog_matching.py:12 opens a file relative to the cwd
    try:
        cm = importlib.import_module("catalog_matching")
    finally:
        os.chdir(cwd)
    logging.disable(logging.CRITICAL)
    u = sys.modules["astropy"].units
    from oracle.make_golden_matching import SampleCluster, make_case
    sample, cat = make_case(100 + seed, 150 + 40 * seed, 120 + 60 * seed, zero_richness=(seed == 2))
    clusters = [SampleCluster(*row) for row in sample]
    matched, idx = cm.compare_clusters(clusters, cat, 0.5 * u.Mpc / u.littleh, richness_plot=False)
    got_m, got_idx = O.compare_clusters(sample, cat, 0.5)
    assert np.array_equal(idx, got_idx) and np.array_equal(matched, got_m)
    assert idx.sum() > 10


def test_parameter_defaults_equal_the_reference_module():
    """fof_b200.params against params.py:16-27 as imported live (names and values), and the C ABI defaults."""
    ref = RR.load_reference()
    from fof_b200 import params as ours
    rp = ref.params
    for name in ("min_redshift", "max_redshift", "linking_length_factor", "richness", "D", "runname"):
        assert getattr(ours, name) == getattr(rp, name), name
    assert float(ours.max_velocity) == float(rp.max_velocity.value)
    assert float(ours.max_radius) == float(rp.max_radius.value)
    assert (ours.cosmo.H0, ours.cosmo.Om0, ours.cosmo.Ode0) == (float(rp.cosmo.H0.value), rp.cosmo.Om0, rp.cosmo.Ode0)
    op = O.Params()
    assert (op.max_velocity, op.linking_length_factor, op.virial_radius, op.richness, op.overdensity) == \
        (float(rp.max_velocity.value), rp.linking_length_factor, float(rp.max_radius.value), rp.richness, float(rp.D))


@pytest.mark.parametrize("n,seed,kw,b,rich,D", [
    (1200, 51, {}, 0.05, 5, -5.0),
    (1500, 52, {"density": 200000.0}, 0.3, 8, 0.0),
    (1800, 53, {"kind": "wide"}, 0.2, 8, 0.5),
    (900, 54, {"cluster_fraction": 0.6, "richness_range": (10, 60)}, 0.1, 5, 2.0),
])
def test_live_reference_parameter_sweep(n, seed, kw, b, rich, D):
    """More live runs of the unmodified reference with other catalogues and parameters, against both oracle modes."""
    import logging
    logging.disable(logging.CRITICAL)
    ref = RR.load_reference()
    df = make_catalog(n, seed=seed, **kw)
    R = RR.run_pipeline(ref, df, richness=rich, overdensity=D, linking_length_factor=b, seed=seed)
    p = O.Params(linking_length_factor=b, richness=rich, overdensity=D)
    for mode in ("fast", "literal"):
        main_arr, _ = O.candidate_center_search(df.values, p, mode)
        assert np.array_equal(main_arr, R["main_arr"])
        g = main_arr[(main_arr[:, 2] >= 0.5) & (main_arr[:, 2] <= 2.52)]
        c = R["candidate_centers"]
        c = c[(c[:, 2] >= 0.5) & (c[:, 2] <= 2.5)]
        np.random.seed(seed)
        cl = O.FoF(g, c, p, mode)
        assert len(cl) == len(R["clusters"])
        for x, (attrs, members, N, Dv) in zip(cl, R["clusters"]):
            assert np.array_equal(x.galaxies, members) and x.N == N
            assert x.D == Dv or (np.isnan(x.D) and np.isnan(Dv))
        cleaned, removed = O.interloper_removal(cl, p)
        assert removed == R["number_removed"]
        assert len(cleaned) == len(R["cleaned"])
        for x, (attrs, members) in zip(cleaned, R["cleaned"]):
            assert np.array_equal(x.galaxies, members)
