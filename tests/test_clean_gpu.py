"""GPU parity of stages 5-6 (interloper_removal, find_masses) and of the facade end to end."""
import numpy as np
import pytest

from fof_b200.mock import make_catalog
from oracle import fof_oracle as O

pytestmark = pytest.mark.gpu
REL = 1e-9  # BASELINE.json north_star: floating-point cluster properties within 1e-9 relative


def _oracle_run(n, seed, **kw):
    arr = make_catalog(n, seed=seed, as_frame=False, **kw)
    return arr, O.run_pipeline(arr, O.Params(), mode="fast", seed=4242)


@pytest.mark.parametrize("n,seed,kw", [(3000, 1, {}), (6000, 12, {}), (20000, 13, {}),
                                         (5000, 14, {"richness_range": (200, 900)})])
def test_facade_pipeline_matches_oracle(lib, n, seed, kw):
    from fof_b200 import fof, mass_analysis, params
    arr, ref = _oracle_run(n, seed, **kw)
    import pandas as pd
    from fof_b200.mock import COLUMNS
    df = pd.DataFrame(arr, columns=COLUMNS)
    main_arr, centres = fof.candidate_center_search(df, max_velocity=params.max_velocity, fname="MOCK")
    assert np.array_equal(main_arr, ref["main_arr"])
    assert np.array_equal(centres, ref["candidate_centers"])
    g = main_arr[(main_arr[:, 2] >= params.min_redshift) & (main_arr[:, 2] <= params.max_redshift + 0.02)]
    c = centres[(centres[:, 2] >= params.min_redshift) & (centres[:, 2] <= params.max_redshift)]
    np.random.seed(4242)
    clusters = fof.FoF(g, c, max_velocity=params.max_velocity, linking_length_factor=params.linking_length_factor,
                       virial_radius=params.max_radius, richness=params.richness, overdensity=params.D)
    assert len(clusters) == len(ref["clusters"])
    for cl, (attrs, members, N, D) in zip(clusters, ref["clusters"]):
        assert np.array_equal(cl.galaxies, members)
        assert np.array_equal(cl.center_attributes[:7], attrs[:7]) and cl.N == N and cl.D == D
    cleaned, removed = fof.interloper_removal(clusters)
    assert removed == ref["number_removed"]
    assert len(cleaned) == len(ref["cleaned"])
    for cl, (attrs, members) in zip(cleaned, ref["cleaned"]):
        assert np.array_equal(cl.galaxies, members), "cleaned members (set or order) differ"
    virial = mass_analysis.find_masses(cleaned, "virial")
    assert len(virial) == len(ref["virial"])
    for v, r in zip(virial, ref["virial"]):
        for key in ("cluster_mass", "mass_err", "vel_disp", "vel_disp_err", "projected_radius", "total_absmag"):
            got = float(getattr(v, key))
            assert abs(got - r[key]) <= REL * abs(r[key]), (key, got, r[key])
        assert v.properties.shape == (11,)


def test_three_sigma_single_cluster(lib):
    from fof_b200 import fof
    arr, ref = _oracle_run(4000, 21)
    attrs, members, N, D = ref["clusters"][0]
    c = O.Cluster(attrs, members.copy())
    want = O.three_sigma(c, members, 10)
    got = fof.three_sigma(c, members, 10)
    assert np.array_equal(got, want)
    want7 = O.three_sigma(c, members, 7)
    assert np.array_equal(fof.three_sigma(c, members, 7), want7)


@pytest.mark.parametrize("n,seed", [(6000, 12), (30000, 31)])
def test_whole_job_equals_staged_facade(lib, n, seed):
    """The device-resident pipeline (what bench.py times) returns what the staged facade returns."""
    from fof_b200 import pipeline
    arr, ref = _oracle_run(n, seed)
    np.random.seed(4242)
    cat = pipeline.find_clusters(arr)
    assert len(cat) == len(ref["virial"])
    ids = arr[:, 6]
    for k, r in enumerate(ref["virial"]):
        assert np.array_equal(ids[cat.members(k)], r["galaxies"][:, 6])
        assert ids[cat.centre_row[k]] == r["center"][6]
        assert cat.N[k] == r["N"] and cat.D[k] == r["D"]
        for j, key in enumerate(("cluster_mass", "mass_err", "vel_disp", "vel_disp_err", "projected_radius",
                                 "total_absmag")):
            assert abs(cat.props[k, j] - r[key]) <= REL * abs(r[key]), (key, cat.props[k, j], r[key])
    lab = cat.labels()
    assert lab.shape == (n,) and (lab >= -1).all()


def test_parameter_sweep_reuses_stage0_and_matches_fresh_runs(lib):
    """pipeline.sweep (fofb_pipeline_rerun) == independent runs for every linking-length factor."""
    from fof_b200 import pipeline
    arr = make_catalog(30000, seed=31, as_frame=False)
    bs = [0.2, 0.05, 0.4, 0.1]
    np.random.seed(11)
    swept = pipeline.sweep(arr, bs)
    np.random.seed(11)
    for b in bs:
        fresh = pipeline.find_clusters(arr, linking_length_factor=b)
        s = swept[b]
        assert np.array_equal(s.offsets, fresh.offsets) and np.array_equal(s.member_rows, fresh.member_rows)
        assert np.array_equal(s.centre_row, fresh.centre_row)
        assert np.array_equal(s.N, fresh.N) and np.array_equal(s.D, fresh.D)
        np.testing.assert_allclose(s.props, fresh.props, rtol=1e-12)
