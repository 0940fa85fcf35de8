"""ctypes binding of libfofb200.so (include/fofb.h).  There is no CPU fallback: if the CUDA library
is missing or no B200 is visible, every compute entry point raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FOFB_LIB") or os.path.join(HERE, "libfofb200.so")  # FOFB_LIB: a developer build (scripts/build_tile_variants.py)


class FofbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libfofb200 error {code}: {message}")
        self.code = code


class Params(C.Structure):
    """fofb_params (include/fofb.h) == params.py:16-27 + inlined magic numbers."""
    _fields_ = [
        ("struct_size", C.c_size_t),
        ("H0", C.c_double), ("Om0", C.c_double), ("Ode0", C.c_double),
        ("max_velocity", C.c_double),
        ("linking_length_factor", C.c_double),
        ("virial_radius", C.c_double),
        ("overdensity", C.c_double),
        ("count_radius", C.c_double),
        ("bcg_fraction", C.c_double),
        ("survey_area", C.c_double),
        ("richness", C.c_int32),
        ("min_window", C.c_int32),
        ("min_count", C.c_int32),
        ("min_virial", C.c_int32),
        ("n_random", C.c_int32),
        ("n_bins", C.c_int32),
        ("grid_cells", C.c_int32),
        ("reserved", C.c_int32),
    ]


class Matrix(C.Structure):
    _fields_ = [("data", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int32), ("mem", C.c_int32),
                ("row_stride", C.c_int64), ("col_stride", C.c_int64)]


_lib = None


def load():
    """Load the shared library (building is ``python -m fof_b200.build`` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m fof_b200.build` (nvcc, sm_100a). "
                          "fof_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    sigs = {
        "fofb_abi_version": (C.c_int, []),
        "fofb_build_info": (C.c_char_p, []),
        "fofb_params_default": (C.c_int, [P(Params)]),
        "fofb_create": (C.c_int, [C.c_int, P(vp)]),
        "fofb_destroy": (None, [vp]),
        "fofb_last_error": (C.c_char_p, [vp]),
        "fofb_synchronize": (C.c_int, [vp]),
        "fofb_last_timing": (C.c_int, [vp, P(dbl), P(i64)]),
        "fofb_profile_enable": (C.c_int, [vp, i32]),
        "fofb_profile_read": (C.c_int, [vp, C.c_char_p, i64]),
        "fofb_comoving_distance": (C.c_int, [P(Params), i64, vp, vp]),
        "fofb_angular_diameter_distance": (C.c_int, [P(Params), i64, vp, vp]),
        "fofb_linear_to_angular_dist": (C.c_int, [P(Params), dbl, i64, vp, vp]),
        "fofb_mean_separation": (C.c_int, [P(Params), dbl, dbl, P(dbl)]),
        "fofb_redshift_to_velocity": (C.c_int, [i64, vp, dbl, vp]),
        "fofb_number_counts": (C.c_int, [vp, P(Params), P(Matrix), vp, i32, P(i64)]),
        "fofb_candidate_order": (C.c_int, [vp, vp, i32, i64]),
        "fofb_fof_run": (C.c_int, [vp, P(Params), P(Matrix), P(Matrix), P(i64), vp]),
        "fofb_fof_finish": (C.c_int, [vp, vp, vp, i32, P(i64)]),
        "fofb_fof_finish_mt": (C.c_int, [vp, vp, P(i32), P(i64)]),
        "fofb_pipeline_finish_mt": (C.c_int, [vp, vp, P(i32), P(i64), P(i64)]),
        "fofb_pipeline_run": (C.c_int, [vp, P(Params), P(Matrix), dbl, dbl, dbl, vp, P(i32), P(i64), P(i64)]),
        "fofb_transitive_labels": (C.c_int, [vp, P(Params), P(Matrix), dbl, vp, i32, P(i64), P(i64)]),
        "fofb_dist_begin": (C.c_int, [vp, P(Params), P(Matrix), vp, vp, vp, P(i64), P(i64)]),
        "fofb_dist_local_inputs": (C.c_int, [vp, vp, vp, vp, vp]),
        "fofb_dist_prepare": (C.c_int, [vp, vp, i64, vp, vp]),
        "fofb_dist_round_evaluate": (C.c_int, [vp]),
        "fofb_dist_state": (C.c_int, [vp, vp, vp, i32]),
        "fofb_dist_round_advance": (C.c_int, [vp, P(i64)]),
        "fofb_set_stream": (C.c_int, [vp, vp]),
        "fofb_dist_overlap_events": (C.c_int, [vp, i64, vp, vp, i64, vp, P(i64)]),
        "fofb_dist_overlap_events_fetch": (C.c_int, [vp, i64, vp, vp, vp]),
        "fofb_dist_overlap_resolve": (C.c_int, [vp, i64, i64, vp, vp, vp, vp]),
        "fofb_dist_state_pack": (C.c_int, [vp, i64, vp, vp, vp, i64]),
        "fofb_dist_state_unpack": (C.c_int, [vp, i64, vp, vp, vp]),
        "fofb_dist_round_advance2": (C.c_int, [vp, vp, i64, P(i64), P(i32)]),
        "fofb_dist_finalize": (C.c_int, [vp, P(i64), P(i64)]),
        "fofb_dist_local_clusters": (C.c_int, [vp, vp, vp, vp]),
        "fofb_dist_overlap": (C.c_int, [vp, i64, vp, vp, vp, vp]),
        "fofb_dist_set_merged": (C.c_int, [vp, vp, vp, i64, vp]),
        "fofb_pipeline_rerun": (C.c_int, [vp, dbl, i32, dbl, vp, P(i32), P(i64), P(i64)]),
        "fofb_bubble_count": (C.c_int, [vp, P(Params), P(Matrix), i64, vp, vp, i32, dbl, vp]),
        "fofb_mt19937_words": (C.c_int, [vp, vp, vp, i64, vp, i32]),
        "fofb_match_catalog": (C.c_int, [vp, P(Params), P(Matrix), P(Matrix), dbl, dbl, vp, vp, i32, vp]),
        "fofb_dist_prefetch_mt": (C.c_int, [vp, vp, i32, i64]),
        "fofb_dist_prefetch_mt_part": (C.c_int, [vp, vp, i32, i64, i32, i32, P(C.c_uint64), P(i64), P(i64)]),
        "fofb_dist_mt_wait": (C.c_int, [vp]),
        "fofb_dist_local_clusters_dev": (C.c_int, [vp, vp, vp, vp]),
        "fofb_dist_overlap_rows": (C.c_int, [vp, i64, vp, vp, i64, vp, vp]),
        "fofb_fof_sizes": (C.c_int, [vp, i32, P(i64), P(i64)]),
        "fofb_fof_fetch": (C.c_int, [vp, i32, vp, vp, vp, vp, vp]),
        "fofb_fof_stats": (C.c_int, [vp, P(i64), P(i64)]),
        "fofb_three_sigma": (C.c_int, [vp, P(Params), i64, vp, P(Matrix), vp, vp, vp]),
        "fofb_virial": (C.c_int, [vp, P(Params), i64, vp, P(Matrix), vp]),
        "fofb_projected_mass": (C.c_int, [vp, P(Params), i64, vp, P(Matrix), vp]),
        "fofb_bootstrap_indices": (C.c_int, [i64, i64, vp]),
        "fofb_pipeline_begin": (C.c_int, [vp, P(Params), P(Matrix), dbl, dbl, dbl, P(i64)]),
        "fofb_pipeline_finish": (C.c_int, [vp, vp, i32, P(i64), P(i64)]),
        "fofb_pipeline_fetch": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
        "fofb_pipeline_stats": (C.c_int, [vp, vp]),
        "fofb_pipeline_fetch_counts": (C.c_int, [vp, vp, i32]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.fofb_abi_version() != 1:
        raise ImportError("libfofb200.so ABI version mismatch; rebuild with `python -m fof_b200.build -f`")
    _lib = lib
    return lib


def default_params(**overrides):
    p = Params()
    load().fofb_params_default(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(f"fofb_params has no field {k!r}")
        setattr(p, k, v)
    return p


def matrix_of(arr):
    """Describe a 2-D float64 numpy array (any strides >= 0) or a torch CUDA tensor to the C ABI."""
    if hasattr(arr, "data_ptr"):  # torch tensor
        if arr.dtype.__str__() != "torch.float64" or arr.dim() != 2:
            raise TypeError("device matrices must be 2-D float64 tensors")
        mem = 1 if arr.is_cuda else 0
        return Matrix(arr.data_ptr(), arr.shape[0], arr.shape[1], mem, arr.stride(0), arr.stride(1)), arr
    a = np.asarray(arr)
    if a.dtype != np.float64 or a.ndim != 2:
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.ndim != 2:
            raise TypeError("expected a 2-D array")
    rs, cs = a.strides[0] // 8, a.strides[1] // 8
    if a.strides[0] % 8 or a.strides[1] % 8 or rs < 0 or cs < 0:
        a = np.ascontiguousarray(a)
        rs, cs = a.strides[0] // 8, a.strides[1] // 8
    return Matrix(a.ctypes.data, a.shape[0], a.shape[1], 0, rs, cs), a


class Handle:
    """One fofb_handle bound to one CUDA device."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.fofb_create(int(device), C.byref(h))
        if rc != 0:
            raise FofbError(rc, self.lib.fofb_last_error(None).decode())
        self.h = h
        self.device = int(device)

    def check(self, rc):
        if rc != 0:
            raise FofbError(rc, self.lib.fofb_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.fofb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def use_stream(self, cuda_stream):
        """Run this handle's kernels on the given CUDA stream (an integer handle such as
        torch.cuda.current_stream().cuda_stream; None: the handle's own stream)."""
        self.check(self.lib.fofb_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def profile(self, on):
        self.check(self.lib.fofb_profile_enable(self.h, int(on)))

    def profile_read(self):
        """{kernel: (total_ms, launches)} accumulated since the last reset."""
        buf = C.create_string_buffer(1 << 16)
        self.check(self.lib.fofb_profile_read(self.h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, ms, cnt = line.split()
            out[name] = (float(ms), int(cnt))
        return out

    def last_timing(self):
        ms, k = C.c_double(), C.c_int64()
        self.lib.fofb_last_timing(self.h, C.byref(ms), C.byref(k))
        return ms.value, k.value

    # ---- stage 0 -----------------------------------------------------------------------------
    def number_counts(self, galaxies, params):
        """N(0.5) per row (int32, host) and the candidate rows sorted by (N desc, row desc)."""
        m, keep = matrix_of(galaxies)
        n = m.rows
        N = np.empty(n, dtype=np.int32)
        ncand = C.c_int64()
        self.check(self.lib.fofb_number_counts(self.h, C.byref(params), C.byref(m), N.ctypes.data, 0, C.byref(ncand)))
        rows = np.empty(ncand.value, dtype=np.int32)
        self.check(self.lib.fofb_candidate_order(self.h, rows.ctypes.data, 0, rows.size))
        del keep
        return N, rows


    # ---- stages 1-4 ----------------------------------------------------------------------------
    def fof_run(self, galaxies, centres, params):
        """Stages 1-3. Returns (number of clusters reaching stage 4, galaxy bbox (4,))."""
        mg, kg = matrix_of(galaxies)
        mc, kc = matrix_of(centres)
        k = C.c_int64()
        bbox = np.empty(4)
        self.check(self.lib.fofb_fof_run(self.h, C.byref(params), C.byref(mg), C.byref(mc), C.byref(k),
                                         bbox.ctypes.data))
        del kg, kc
        return k.value, bbox

    def fof_finish(self, rand_ra, rand_dec):
        """Stage 4 with caller-drawn uniform points (cluster-major, n_random per cluster)."""
        ra = np.ascontiguousarray(rand_ra, dtype=np.float64)
        dec = np.ascontiguousarray(rand_dec, dtype=np.float64)
        k = C.c_int64()
        self.check(self.lib.fofb_fof_finish(self.h, ra.ctypes.data, dec.ctypes.data, 0, C.byref(k)))
        return k.value

    @staticmethod
    def _numpy_mt_state():
        name, key, pos, has_gauss, cached = np.random.get_state()
        if name != "MT19937":
            raise RuntimeError("numpy's legacy generator is not MT19937")
        return np.ascontiguousarray(key, dtype=np.uint32).copy(), C.c_int32(int(pos)), has_gauss, cached

    def fof_finish_np_random(self):
        """Stage 4 with the points drawn on the device from np.random's own MT19937 state, which is advanced
        exactly as helper.py:197-202 would advance it."""
        key, pos, has_gauss, cached = self._numpy_mt_state()
        k = C.c_int64()
        self.check(self.lib.fofb_fof_finish_mt(self.h, key.ctypes.data, C.byref(pos), C.byref(k)))
        np.random.set_state(("MT19937", key, pos.value, has_gauss, cached))
        return k.value

    def pipeline_finish_np_random(self):
        key, pos, has_gauss, cached = self._numpy_mt_state()
        k, t = C.c_int64(), C.c_int64()
        self.check(self.lib.fofb_pipeline_finish_mt(self.h, key.ctypes.data, C.byref(pos), C.byref(k), C.byref(t)))
        np.random.set_state(("MT19937", key, pos.value, has_gauss, cached))
        return k.value, t.value

    def pipeline_finish_mt(self, key, pos):
        """Like pipeline_finish_np_random with an explicit MT19937 state; returns (K, total, key, pos)."""
        key = np.ascontiguousarray(key, dtype=np.uint32).copy()
        cpos = C.c_int32(int(pos))
        k, t = C.c_int64(), C.c_int64()
        self.check(self.lib.fofb_pipeline_finish_mt(self.h, key.ctypes.data, C.byref(cpos), C.byref(k), C.byref(t)))
        return k.value, t.value, key, cpos.value

    def pipeline_run_np_random(self, galaxies, params, zmin, zmax_gal, zmax_cen):
        """The whole path in one C call; np.random's MT19937 state is consumed on the device and written back."""
        m, keep = matrix_of(galaxies)
        key, pos, has_gauss, cached = self._numpy_mt_state()
        k, t = C.c_int64(), C.c_int64()
        self.check(self.lib.fofb_pipeline_run(self.h, C.byref(params), C.byref(m), zmin, zmax_gal, zmax_cen,
                                              key.ctypes.data, C.byref(pos), C.byref(k), C.byref(t)))
        np.random.set_state(("MT19937", key, pos.value, has_gauss, cached))
        del keep
        return k.value, t.value

    def pipeline_rerun_np_random(self, linking_length_factor, richness, overdensity):
        """Parameter sweep step on the catalogue of the last pipeline_run (stage 0 and the catalogue are reused)."""
        key, pos, has_gauss, cached = self._numpy_mt_state()
        k, t = C.c_int64(), C.c_int64()
        self.check(self.lib.fofb_pipeline_rerun(self.h, float(linking_length_factor), int(richness), float(overdensity),
                                                key.ctypes.data, C.byref(pos), C.byref(k), C.byref(t)))
        np.random.set_state(("MT19937", key, pos.value, has_gauss, cached))
        return k.value, t.value

    def fof_fetch(self, which=3):
        """(offsets int64[K+1], members int32[total], centre int32[K], N float64[K], D float64[K])."""
        k, t = C.c_int64(), C.c_int64()
        self.check(self.lib.fofb_fof_sizes(self.h, which, C.byref(k), C.byref(t)))
        K, total = k.value, t.value
        off = np.zeros(K + 1, dtype=np.int64)
        mem = np.empty(total, dtype=np.int32)
        cen = np.empty(K, dtype=np.int32)
        N = np.empty(K)
        D = np.empty(K)
        self.check(self.lib.fofb_fof_fetch(self.h, which, off.ctypes.data, mem.ctypes.data, cen.ctypes.data,
                                           N.ctypes.data, D.ctypes.data))
        return off, mem, cen, N, D

    # ---- stages 5-6 ----------------------------------------------------------------------------
    def three_sigma(self, offsets, members, centres, params):
        """(out_offsets int64[K+1], out_index int32[kept]) -- positions inside each cluster."""
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        cen = np.ascontiguousarray(centres, dtype=np.float64)
        K = len(off) - 1
        m, keep = matrix_of(members)
        out_off = np.zeros(K + 1, dtype=np.int64)
        out_idx = np.empty(max(int(off[-1]), 1), dtype=np.int32)
        self.check(self.lib.fofb_three_sigma(self.h, C.byref(params), K, off.ctypes.data, C.byref(m), cen.ctypes.data,
                                             out_off.ctypes.data, out_idx.ctypes.data))
        del keep
        return out_off, out_idx[:out_off[K]]

    def virial(self, offsets, members, params):
        """(K, 8): mass, mass_err, vel_disp, vel_disp_err, radius, total_absmag, mean z, n."""
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        K = len(off) - 1
        m, keep = matrix_of(members)
        out = np.empty((K, 8))
        self.check(self.lib.fofb_virial(self.h, C.byref(params), K, off.ctypes.data, C.byref(m), out.ctypes.data))
        del keep
        return out

    # ---- whole job ------------------------------------------------------------------------------
    def pipeline_begin(self, galaxies, params, zmin, zmax_gal, zmax_cen):
        m, keep = matrix_of(galaxies)
        k = C.c_int64()
        self.check(self.lib.fofb_pipeline_begin(self.h, C.byref(params), C.byref(m), zmin, zmax_gal, zmax_cen,
                                                C.byref(k)))
        del keep
        return k.value

    def pipeline_finish(self, u):
        """u: host ndarray (K, 2, n_random) of [0, 1) samples, or a (device_ptr, ) tuple for resident samples."""
        k, t = C.c_int64(), C.c_int64()
        if isinstance(u, tuple):
            ptr, mem = u[0], 1
        else:
            u = np.ascontiguousarray(u, dtype=np.float64)
            ptr, mem = u.ctypes.data, 0
        self.check(self.lib.fofb_pipeline_finish(self.h, ptr, mem, C.byref(k), C.byref(t)))
        return k.value, t.value

    def pipeline_fetch(self, K, total):
        off = np.zeros(K + 1, dtype=np.int64)
        rows = np.empty(total, dtype=np.int32)
        cen = np.empty(K, dtype=np.int32)
        N = np.empty(K)
        D = np.empty(K)
        props = np.empty((K, 8))
        self.check(self.lib.fofb_pipeline_fetch(self.h, off.ctypes.data, rows.ctypes.data, cen.ctypes.data,
                                                N.ctypes.data, D.ctypes.data, props.ctypes.data))
        return off, rows, cen, N, D, props

    def pipeline_fetch_counts(self, n_rows):
        """N(0.5) of every input row of the last pipeline run (int32, host)."""
        N = np.empty(int(n_rows), dtype=np.int32)
        self.check(self.lib.fofb_pipeline_fetch_counts(self.h, N.ctypes.data, 0))
        return N

    def pipeline_fetch_torch(self, K, total, device):
        """The same as pipeline_fetch into torch tensors on ``device`` (no host round trip)."""
        import torch
        off = torch.zeros(K + 1, dtype=torch.int64, device=device)
        rows = torch.empty(max(total, 1), dtype=torch.int32, device=device)[:total]
        cen = torch.empty(max(K, 1), dtype=torch.int32, device=device)[:K]
        N = torch.empty(max(K, 1), dtype=torch.float64, device=device)[:K]
        D = torch.empty(max(K, 1), dtype=torch.float64, device=device)[:K]
        props = torch.empty((max(K, 1), 8), dtype=torch.float64, device=device)[:K]
        self.check(self.lib.fofb_pipeline_fetch(self.h, off.data_ptr(), rows.data_ptr(), cen.data_ptr(), N.data_ptr(),
                                                D.data_ptr(), props.data_ptr()))
        return off, rows, cen, N, D, props

    def pipeline_stats(self):
        s = np.zeros(7, dtype=np.int64)
        self.lib.fofb_pipeline_stats(self.h, s.ctypes.data)
        return dict(zip(("galaxies_after_zcut", "candidate_centres", "stage1_clusters", "stage1_members",
                         "priority_rounds", "centres_evaluated", "stage4_clusters"), s.tolist()))

    def mt19937_words(self, n_words, state=None):
        """The next n_words raw 32-bit outputs of numpy's legacy generator, drawn on the device.  state: a
        np.random.get_state() tuple (default: the global generator, which is advanced).  Returns (words, new_state)."""
        use_global = state is None
        st = np.random.get_state() if use_global else state
        key = np.ascontiguousarray(st[1], dtype=np.uint32).copy()
        pos = C.c_int32(int(st[2]))
        out = np.empty(int(n_words), dtype=np.uint32)
        self.check(self.lib.fofb_mt19937_words(self.h, key.ctypes.data, C.byref(pos), int(n_words), out.ctypes.data, 0))
        new_state = ("MT19937", key, pos.value, st[3], st[4])  # a pending cached Gaussian is not the uniforms' business
        if use_global:
            np.random.set_state(new_state)
        return out, new_state

    def match_catalog(self, sample, catalogue, params, matching_distance_mpc_h, z_cut=0.1):
        """(match_row int32[C], match_sep_deg float64[C], n_matched) -- see include/fofb.h fofb_match_catalog."""
        ms, keep_s = matrix_of(sample)
        mc, keep_c = matrix_of(catalogue)
        rows = np.empty(mc.rows, dtype=np.int32)
        sep = np.empty(mc.rows, dtype=np.float64)
        n = C.c_int64()
        self.check(self.lib.fofb_match_catalog(self.h, C.byref(params), C.byref(ms), C.byref(mc), float(matching_distance_mpc_h),
                                               float(z_cut), rows.ctypes.data, sep.ctypes.data, 0, C.byref(n)))
        del keep_s, keep_c
        return rows, sep, n.value

    def transitive_labels(self, galaxies, params, linking_length_mpc_h, count_links=True):
        """(labels int32[n] = smallest row of each row's component, number of groups, number of links or -1)."""
        m, keep = matrix_of(galaxies)
        lab = np.empty(m.rows, dtype=np.int32)
        g, l = C.c_int64(), C.c_int64(-1)
        self.check(self.lib.fofb_transitive_labels(self.h, C.byref(params), C.byref(m), float(linking_length_mpc_h),
                                                   lab.ctypes.data, 0, C.byref(g), C.byref(l) if count_links else None))
        del keep
        return lab, g.value, l.value

    def projected(self, offsets, members, params):
        """(K, 2): projected mass [Msun/h], mass_err."""
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        K = len(off) - 1
        m, keep = matrix_of(members)
        out = np.empty((K, 2))
        self.check(self.lib.fofb_projected_mass(self.h, C.byref(params), K, off.ctypes.data, C.byref(m), out.ctypes.data))
        del keep
        return out

    def bubble_count(self, points, centres, radius_deg, params, window_z=None):
        """Counts of GriSPy(points[window]).bubble_neighbors(centres, radius) (haversine, degrees)."""
        m, keep = matrix_of(points)
        cen = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, 2)
        rad = np.ascontiguousarray(np.broadcast_to(np.asarray(radius_deg, dtype=np.float64), (len(cen),)))
        out = np.empty(len(cen), dtype=np.int32)
        self.check(self.lib.fofb_bubble_count(self.h, C.byref(params), C.byref(m), len(cen), cen.ctypes.data, rad.ctypes.data,
                                              0 if window_z is None else 1, 0.0 if window_z is None else float(window_z),
                                              out.ctypes.data))
        del keep
        return out

    def fof_stats(self):
        r, e = C.c_int64(), C.c_int64()
        self.lib.fofb_fof_stats(self.h, C.byref(r), C.byref(e))
        return r.value, e.value


_default_handles = {}


def get_handle(device=0):
    h = _default_handles.get(device)
    if h is None:
        h = _default_handles[device] = Handle(device)
    return h
