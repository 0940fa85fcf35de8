"""fof_b200: the cluster-finding hot path of kencx/FoF on a B200, behind the reference's Python API.

    from fof_b200.fof import candidate_center_search, FoF, interloper_removal, three_sigma
    from fof_b200.mass_analysis import find_masses
    from fof_b200 import params
    from fof_b200.pipeline import find_clusters, sweep          # the same sequence as one device-resident job
    from fof_b200.transitive import friends_of_friends          # union-find mode (not a reference code path)

Everything is computed by hand-written CUDA (libfofb200.so, built by ``python -m fof_b200.build``).
"""
__version__ = "0.1.0"
__all__ = ["fof", "mass_analysis", "cluster", "params", "helper", "pipeline", "transitive", "distributed", "mock", "units"]
