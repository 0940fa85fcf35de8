"""Run the UNMODIFIED reference files from /root/reference/FoF against the third-party shims.

TEST INFRASTRUCTURE ONLY.  Works only in the build container (where /root/reference is mounted);
nothing that runs on the GPU box imports this module.  It is used by ``oracle/make_golden.py`` to
generate the fixtures under tests/golden/ and by tests that cross-check ``oracle/fof_oracle.py``
(our restatement) against the reference's own control flow.

The reference's own source is executed as-is: fof.py, helper.py, cluster.py, params.py,
mass_analysis.py, analysis/methods.py, analysis/mass_estimator.py.  Only astropy, grispy and
matplotlib (absent from this image, un-vendored and un-pinned upstream) are replaced by the
restatements under oracle/shims/.
"""
import contextlib
import importlib
import io
import os
import sys
import tempfile

REFERENCE_DIR = "/root/reference/FoF"
SHIMS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_MODS = ("params", "cluster", "helper", "fof", "mass_analysis", "analysis", "analysis.methods",
         "analysis.mass_estimator")


def available():
    return os.path.isdir(REFERENCE_DIR)


class Reference:
    """Namespace holding the imported reference modules."""


def load_reference(quiet=True):
    if not available():
        raise RuntimeError("reference tree not mounted at /root/reference")
    for name in ("astropy", "grispy", "matplotlib"):
        mod = sys.modules.get(name)
        if mod is not None and SHIMS_DIR not in getattr(mod, "__file__", ""):
            raise RuntimeError(f"a real {name} is already imported; refusing to shadow it")
    for p in (REFERENCE_DIR, SHIMS_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    scratch = tempfile.mkdtemp(prefix="fof_ref_")  # params.py:38 creates directories on import
    cwd = os.getcwd()
    os.chdir(scratch)
    ref = Reference()
    # params.py:34-38 builds a Windows "\\"-separated path; on POSIX its dirname is '' and
    # os.makedirs('') raises.  The file is executed unmodified; only that one call is tolerated.
    real_makedirs = os.makedirs
    os.makedirs = lambda name, *a, **k: None if name == "" else real_makedirs(name, *a, **k)
    try:
        sink = io.StringIO()
        with contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext():
            for name in _MODS:
                setattr(ref, name.replace(".", "_"), importlib.import_module(name))
    finally:
        os.makedirs = real_makedirs
        os.chdir(cwd)
    ref.scratch = scratch
    return ref


def run_pipeline(ref, df, richness=None, overdensity=None, linking_length_factor=None, seed=12345,
                 quiet=True):
    """pipeline.py:62-64,80-99,113,173 call order on a DataFrame; returns every stage artefact."""
    import numpy as np

    p = ref.params
    richness = p.richness if richness is None else richness
    overdensity = p.D if overdensity is None else overdensity
    b = p.linking_length_factor if linking_length_factor is None else linking_length_factor
    out = {}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext():
        main_arr, centers = ref.fof.candidate_center_search(
            df, max_velocity=p.max_velocity, fname="MOCK", path=ref.scratch + os.sep)
        out["main_arr"], out["candidate_centers"] = main_arr.copy(), centers.copy()
        g = main_arr[(main_arr[:, 2] >= p.min_redshift) & (main_arr[:, 2] <= p.max_redshift + 0.02)]
        c = centers[(centers[:, 2] >= p.min_redshift) & (centers[:, 2] <= p.max_redshift)]
        np.random.seed(seed)  # pin D3: center_overdensity draws from the global RNG
        clusters = ref.fof.FoF(g, c, max_velocity=p.max_velocity, linking_length_factor=b,
                               virial_radius=p.max_radius, richness=richness, overdensity=overdensity)
        out["clusters"] = [(cl.center_attributes.copy(), cl.galaxies.copy(), float(cl.N), float(cl.D))
                           for cl in clusters]
        # interloper_removal is in fof.py (pipeline.py:113 forgets the import); it reads the
        # module-global `richness` star-imported from params (fof.py:514)
        ref.fof.richness = richness
        cleaned, removed = ref.fof.interloper_removal(clusters)
        out["cleaned"] = [(cl.center_attributes.copy(), cl.galaxies.copy()) for cl in cleaned]
        out["number_removed"] = list(removed)
        out["cleaned_objects"] = cleaned   # the reference's own Cluster instances (oracle/make_golden_io.py pickles them)
        virial = ref.mass_analysis.find_masses(cleaned, "virial")
        out["virial"] = [
            dict(center=v.center_attributes.copy(), galaxies=v.galaxies.copy(),
                 cluster_mass=float(v.cluster_mass.value), mass_err=float(v.mass_err.value),
                 vel_disp=float(v.vel_disp.value), vel_disp_err=float(v.vel_disp_err.value),
                 projected_radius=float(v.projected_radius.value), total_absmag=float(v.total_absmag),
                 # cluster.py:111-118 with the reference's own cosmology object, delta = 200 and 500
                 r200=float(v.r200(p.cosmo, 200).value), r500=float(v.r200(p.cosmo, 500).value))
            for v in virial]
        proj = ref.mass_analysis.find_masses(cleaned, "projected")
        out["projected"] = [(float(v.cluster_mass.value), float(getattr(v.mass_err, "value", v.mass_err))) for v in proj]
    return out
