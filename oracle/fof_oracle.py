"""CPU oracle: a numpy restatement of the reference FoF cluster-finding hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product (fof_b200/) never does.

PARITY STATUS: the restatement of the reference's OWN code (fof.py, helper.py, cluster.py,
mass_analysis.py, analysis/*.py) is pinned against the unmodified reference files executed in the
build container (oracle/run_reference.py -> tests/golden/, tests/test_oracle_vs_reference.py).
The arithmetic that lives in the un-vendored, un-pinned third-party packages GriSPy and astropy is
restated from their published algorithms and is PARITY UNPINNED: the reference
holds no tests, golden vectors or fixtures for it.

Pinned orderings: D1 equal-N candidates in descending original index
(``argsort(kind="stable")[::-1]``); D2 GriSPy neighbour order = cell-major (dec cell outer, RA cell
inner), ascending index inside a cell; D3 the caller seeds numpy's legacy global RNG right before
``FoF`` and the 2x300 uniforms per cluster are drawn from it in cluster order; D4 exact ties in the
3-sigma deviation sort (fof.py:456-458) keep their order.

Two execution modes with identical results:
  mode="literal"  the reference's control flow: a full-catalogue velocity mask and a fresh 64x64 grid
                  per centre (O(n^2); this is the timed CPU baseline);
  mode="fast"     z-sorted windows + a k-d tree for candidate generation (finishes at 1M galaxies).
Column layout (pipeline.py:51-53 + fof.py:61): 0 ra 1 dec 2 z 3 z_lower 4 z_upper 5 RMag 6 ID 7 N.
"""
from dataclasses import dataclass
import math

import numpy as np
from scipy.integrate import quad
from scipy.special import hyp2f1

# ------------------------------------------------------------------------------------------------
# constants and cosmology (params.py:16; analysis/methods.py:9)
# ------------------------------------------------------------------------------------------------
C_KMS = 299792.458
H0 = 70.0
OM0 = 0.3
ODE0 = 0.7
LITTLE_H = H0 / 100.0
D_H = C_KMS / H0  # Mpc
N_CELLS = 64
# (km/s)^2 Mpc / G in solar masses: CODATA-2018 G, IAU-2015 GM_sun, au -> pc
_PC_M = 149597870700.0 * 648000.0 / math.pi
_MSUN_KG = 1.3271244e20 / 6.6743e-11
MASS_UNIT = 1.0e6 * (1.0e6 * _PC_M) / 6.6743e-11 / _MSUN_KG


@dataclass
class Params:
    """params.py:23-27 plus the magic numbers inlined in fof.py / helper.py."""
    max_velocity: float = 2000.0          # km/s          params.py:23
    linking_length_factor: float = 0.2    # b             params.py:24 (FoF() default is 0.1, fof.py:104)
    virial_radius: float = 1.5            # Mpc/h         params.py:25
    richness: int = 12                    #               params.py:26
    overdensity: float = 2.0              # D             params.py:27
    min_window: int = 8                   # fof.py:71
    min_count: int = 8                    # fof.py:77,386
    min_virial: int = 12                  # fof.py:185
    count_radius: float = 0.5             # Mpc/h         fof.py:73,378; helper.py:212
    bcg_fraction: float = 0.25            # fof.py:334
    n_random: int = 300                   # helper.py:196
    survey_area: float = 1.7              # fof.py:192 (used as a radius in deg, Q8)
    n_bins: int = 10                      # fof.py:506


def efunc(z):
    zp1 = 1.0 + np.asarray(z, dtype=float)
    return np.sqrt(zp1 ** 2 * (OM0 * zp1 + 0.0) + ODE0)


def _T(x):
    return 2 * np.sqrt(x) * hyp2f1(1.0 / 6, 1.0 / 2, 7.0 / 6, -(x ** 3))


def comoving_distance(z):
    """D_C [Mpc]; astropy's closed form for flat, radiation-free LCDM."""
    s = ((1 - OM0) / OM0) ** (1.0 / 3)
    pref = D_H / np.sqrt(s * OM0)
    z = np.asarray(z, dtype=float)
    return pref * (_T(s / 1.0) - _T(s / (z + 1.0)))


def angular_diameter_distance(z):
    z = np.asarray(z, dtype=float)
    return comoving_distance(z) / (1.0 + z)


def proper_to_deg(d_mpc, z):
    """analysis/methods.py:30-32 for a length already in plain Mpc."""
    return np.rad2deg(d_mpc / angular_diameter_distance(z))


def linear_to_angular_dist(d_mpc_h, z):
    """analysis/methods.py:12-32: proper h^-1 Mpc -> degrees."""
    return proper_to_deg(d_mpc_h / LITTLE_H, z)


def differential_comoving_volume(z):
    return D_H * comoving_distance(z) ** 2 / efunc(z)


def mean_separation(n, z, max_velocity, survey_area):
    """analysis/methods.py:35-65 (max_dist argument unused there, Q8). Returns Mpc."""
    dz = max_velocity / C_KMS
    solid_angle = math.pi * (survey_area * (math.pi / 180.0)) ** 2
    V = quad(lambda x: float(differential_comoving_volume(x)), z, z + dz)[0] * solid_angle
    return (n / V) ** (-1.0 / 3)


def redshift_to_velocity(z, center_z):
    """analysis/methods.py:68-87: two products, a subtraction and a division, in that association."""
    return (C_KMS * z - C_KMS * center_z) / (1 + center_z)


def haversine_deg(c_ra, c_dec, ra, dec):
    """GriSPy haversine metric, degrees in / degrees out."""
    lon1 = np.deg2rad(c_ra)
    lat1 = np.deg2rad(c_dec)
    lon2 = np.deg2rad(ra)
    lat2 = np.deg2rad(dec)
    sdlon = np.sin((lon2 - lon1) / 2.0)
    sdlat = np.sin((lat2 - lat1) / 2.0)
    num1 = sdlat ** 2
    num2 = np.cos(lat1) * np.cos(lat2) * sdlon ** 2
    return np.rad2deg(2 * np.arcsin(np.sqrt(num1 + num2)))


def vincenty_rad(lon1, lat1, lon2, lat2):
    """astropy angular_separation (radians in/out), used by SkyCoord.separation (App. B.2)."""
    sdlon = np.sin(lon2 - lon1)
    cdlon = np.cos(lon2 - lon1)
    slat1, slat2 = np.sin(lat1), np.sin(lat2)
    clat1, clat2 = np.cos(lat1), np.cos(lat2)
    num1 = clat2 * sdlon
    num2 = clat1 * slat2 - slat1 * clat2 * cdlon
    den = slat1 * slat2 + clat1 * clat2 * cdlon
    return np.arctan2(np.hypot(num1, num2), den)


# ------------------------------------------------------------------------------------------------
# GriSPy grid semantics as pure functions
# ------------------------------------------------------------------------------------------------
def grid_edges(vmin, vmax):
    """(bins[0], bins[-1]) of linspace(min-1e-6, max+1e-6, 65)."""
    return vmin - 1.0e-6, vmax + 1.0e-6


def digitize(x, b0, bN):
    return (N_CELLS * (np.asarray(x, dtype=float) - b0) / (bN - b0)).astype(np.int64)


def _clamp(i):
    return np.minimum(np.maximum(i, 0), N_CELLS - 1)


def grid_query(ra, dec, bbox, c_ra, c_dec, r, out_of_field_all=None):
    """bubble_neighbors for ONE centre over points (ra, dec) that belong to a grid built on a point
    set with bounding box ``bbox`` = (ra_min, ra_max, dec_min, dec_max).

    ``ra, dec`` must be listed in ascending index of that point set (a subset is fine: leaving out
    points that fail the test does not change the answer).  Returns positions into ra/dec in
    GriSPy's return order.  ``out_of_field_all``: result of the all-centres test for multi-centre
    calls; None => single-centre call, computed here.
    """
    bx0, bxN = grid_edges(bbox[0], bbox[1])
    by0, byN = grid_edges(bbox[2], bbox[3])
    if out_of_field_all is None:
        out_of_field_all = (c_ra - r > bxN) or (c_ra + r < bx0) or (c_dec - r > byN) or (c_dec + r < by0)
    if out_of_field_all or len(ra) == 0:
        return np.zeros(0, dtype=np.int64)
    ix_lo, ix_hi = _clamp(digitize(c_ra - r, bx0, bxN)), _clamp(digitize(c_ra + r, bx0, bxN))
    iy_lo, iy_hi = _clamp(digitize(c_dec - r, by0, byN)), _clamp(digitize(c_dec + r, by0, byN))
    ix = digitize(ra, bx0, bxN)
    iy = digitize(dec, by0, byN)
    inbox = (ix >= ix_lo) & (ix <= ix_hi) & (iy >= iy_lo) & (iy <= iy_hi)
    pos = np.nonzero(inbox)[0]
    if len(pos) == 0:
        return pos
    d = haversine_deg(c_ra, c_dec, ra[pos], dec[pos])
    pos = pos[d <= r]
    # cell-major order: np.meshgrid('xy') flattening => dec-cell outer, RA-cell inner; ascending
    # index inside a cell (stable sort keeps it)
    key = iy[pos] * N_CELLS + ix[pos]
    return pos[np.argsort(key, kind="stable")]


def out_of_field(bbox, c_ra, c_dec, r):
    bx0, bxN = grid_edges(bbox[0], bbox[1])
    by0, byN = grid_edges(bbox[2], bbox[3])
    c_ra = np.asarray(c_ra)
    c_dec = np.asarray(c_dec)
    return (c_ra - r > bxN) | (c_ra + r < bx0) | (c_dec - r > byN) | (c_dec + r < by0)


def bbox_of(ra, dec):
    return (ra.min(), ra.max(), dec.min(), dec.max())


# ------------------------------------------------------------------------------------------------
# z-sorted window index used by mode="fast"
# ------------------------------------------------------------------------------------------------
class WindowIndex:
    """Velocity windows as contiguous ranges of the z-sorted catalogue, plus exact range min/max of
    ra/dec over such ranges (block prefix/suffix + sparse table over blocks) and a k-d tree."""
    BLOCK = 256

    def __init__(self, data, max_velocity):
        from scipy.spatial import cKDTree
        self.data = data
        self.vmax = max_velocity
        self.order = np.argsort(data[:, 2], kind="stable")
        self.zs = data[self.order, 2]
        self.rank = np.empty(len(data), dtype=np.int64)
        self.rank[self.order] = np.arange(len(data))
        n = len(data)
        B = self.BLOCK
        nb = (n + B - 1) // B
        self.tables = []
        for col, fn, pad in ((0, np.minimum, np.inf), (0, np.maximum, -np.inf),
                             (1, np.minimum, np.inf), (1, np.maximum, -np.inf)):
            v = np.full(nb * B, pad)
            v[:n] = data[self.order, col]
            blk = v.reshape(nb, B)
            pre = fn.accumulate(blk, axis=1).reshape(-1)
            suf = fn.accumulate(blk[:, ::-1], axis=1)[:, ::-1].reshape(-1)
            levels = [fn.reduce(blk, axis=1)]
            k = 1
            while 2 * k <= nb:
                prev = levels[-1]
                levels.append(fn(prev[:-k], prev[k:]))
                k *= 2
            self.tables.append((fn, pre, suf, levels))
        dec = data[:, 1]
        self.cosmin = np.cos(np.deg2rad(min(89.0, np.abs(dec).max() + 1.0)))
        self.tree = cKDTree(data[:, :2])

    def window(self, zc):
        """[lo, hi) in z-sorted order of {j : |v(z_j; zc)| <= vmax}; exact w.r.t. the predicate."""
        dz = self.vmax * (1 + zc) / C_KMS
        lo = int(np.searchsorted(self.zs, zc - dz * (1 + 1e-9), side="left"))
        hi = int(np.searchsorted(self.zs, zc + dz * (1 + 1e-9), side="right"))
        lo = max(lo - 2, 0)
        hi = min(hi + 2, len(self.zs))
        ok = np.abs(redshift_to_velocity(self.zs[lo:hi], zc)) <= self.vmax
        nz = np.nonzero(ok)[0]
        if len(nz) == 0:
            return lo, lo
        return lo + int(nz[0]), lo + int(nz[-1]) + 1

    def _range(self, t, lo, hi):
        fn, pre, suf, levels = self.tables[t]
        B = self.BLOCK
        b_lo, b_hi = lo // B, (hi - 1) // B
        if b_lo == b_hi:
            return fn.reduce(self.data[self.order[lo:hi], 0 if t < 2 else 1])
        val = fn(suf[lo], pre[hi - 1])
        a, b = b_lo + 1, b_hi  # full blocks [a, b)
        if b > a:
            k = (b - a).bit_length() - 1
            val = fn(val, fn(levels[k][a], levels[k][b - (1 << k)]))
        return val

    def bbox(self, lo, hi):
        return tuple(self._range(t, lo, hi) for t in range(4))

    def candidates(self, c_ra, c_dec, r, lo, hi):
        """Ascending row indices of window members that could lie within r (superset)."""
        reach = r / self.cosmin * 1.001 + 1e-9
        idx = np.asarray(self.tree.query_ball_point([c_ra, c_dec], reach * math.sqrt(2.0), p=2.0), dtype=np.int64)
        if len(idx) == 0:
            return idx
        rk = self.rank[idx]
        idx = idx[(rk >= lo) & (rk < hi)]
        idx.sort()
        return idx


# ------------------------------------------------------------------------------------------------
# Cluster records (cluster.py:7-136)
# ------------------------------------------------------------------------------------------------
class Cluster:
    """cluster.py:7-85 (plain floats instead of astropy quantities)."""

    def __init__(self, center, galaxies):
        self.ra, self.dec, self.z = center[0], center[1], center[2]
        self.z_unc = [center[3], center[4]]
        self.bcg_absMag = center[5]
        self.gal_id = center[6]
        self.N = center[-1]
        self.galaxies = galaxies
        self.D = 0
        self.flag_poor = False

    @property
    def richness(self):
        return len(self.galaxies)

    @property
    def center_attributes(self):
        return np.array([self.ra, self.dec, self.z, self.z_unc[0], self.z_unc[1], self.bcg_absMag,
                         self.gal_id, self.N])

    def remove_overlap(self, other):
        """cluster.py:65-82: (self_galaxies, other_galaxies); the loser is emptied."""
        if self.N > other.N:
            return self.galaxies, np.array([])
        if self.N < other.N:
            return np.array([]), other.galaxies
        if abs(self.bcg_absMag) > abs(other.bcg_absMag):
            return self.galaxies, np.array([])
        return np.array([]), other.galaxies


# ------------------------------------------------------------------------------------------------
# stage 0: candidate_center_search (fof.py:24-95; helper.py:149-188)
# ------------------------------------------------------------------------------------------------
def number_counts_literal(galaxy_arr, p, rows=None):
    """fof.py:65-74 exactly as written: full-array mask and a fresh grid per galaxy."""
    n = len(galaxy_arr)
    N = np.zeros(n)
    ra, dec, z = galaxy_arr[:, 0], galaxy_arr[:, 1], galaxy_arr[:, 2]
    for i in (range(n) if rows is None else rows):
        mask = np.abs(redshift_to_velocity(z, z[i])) <= p.max_velocity
        if mask.sum() >= p.min_window:
            wra, wdec = ra[mask], dec[mask]
            theta = float(linear_to_angular_dist(p.count_radius, z[i]))
            N[i] = len(grid_query(wra, wdec, bbox_of(wra, wdec), ra[i], dec[i], theta))
    return N


def number_counts_fast(galaxy_arr, p, widx=None):
    """Same N vector through z-sorted windows and k-d tree candidate pairs."""
    n = len(galaxy_arr)
    widx = widx or WindowIndex(galaxy_arr, p.max_velocity)
    ra, dec, z = galaxy_arr[:, 0], galaxy_arr[:, 1], galaxy_arr[:, 2]
    theta = linear_to_angular_dist(p.count_radius, z)
    # windows per galaxy (vectorised bracket + exact fix-up)
    dz = p.max_velocity * (1 + z) / C_KMS
    lo = np.searchsorted(widx.zs, z - dz * (1 + 1e-9), side="left")
    hi = np.searchsorted(widx.zs, z + dz * (1 + 1e-9), side="right")
    zs = widx.zs
    for _ in range(4):  # shrink to the exact predicate (a couple of ulps at most)
        bad = (lo < hi) & ~(np.abs(redshift_to_velocity(zs[np.minimum(lo, n - 1)], z)) <= p.max_velocity)
        lo = lo + bad
        bad = (lo < hi) & ~(np.abs(redshift_to_velocity(zs[np.maximum(hi - 1, 0)], z)) <= p.max_velocity)
        hi = hi - bad
    for _ in range(4):  # and grow if the bracket was too tight
        ok = (lo > 0) & (np.abs(redshift_to_velocity(zs[np.maximum(lo - 1, 0)], z)) <= p.max_velocity)
        lo = lo - ok
        ok = (hi < n) & (np.abs(redshift_to_velocity(zs[np.minimum(hi, n - 1)], z)) <= p.max_velocity)
        hi = hi + ok
    wcount = hi - lo
    # candidate directed pairs (i <- j): unordered pairs from the tree plus self pairs
    rmax = float(theta.max())
    reach = rmax / widx.cosmin * 1.001 + 1e-9
    pairs = widx.tree.query_pairs(reach * math.sqrt(2.0), p=2.0, output_type="ndarray")
    qi = np.concatenate([pairs[:, 0], pairs[:, 1], np.arange(n)])
    pj = np.concatenate([pairs[:, 1], pairs[:, 0], np.arange(n)])
    rk = widx.rank[pj]
    keep = (rk >= lo[qi]) & (rk < hi[qi]) & (wcount[qi] >= p.min_window)
    qi, pj = qi[keep], pj[keep]
    d = haversine_deg(ra[qi], dec[qi], ra[pj], dec[pj])
    keep = d <= theta[qi]
    qi, pj = qi[keep], pj[keep]
    # GriSPy cell-box predicate with the per-query grid (bbox of the query's window)
    uq = np.unique(qi)
    if len(uq) == 0:
        return np.zeros(n)
    bb = np.array([widx.bbox(int(lo[i]), int(hi[i])) for i in uq])
    bmap = np.full(n, -1, dtype=np.int64)
    bmap[uq] = np.arange(len(uq))
    b = bb[bmap[qi]]
    bx0, bxN = b[:, 0] - 1e-6, b[:, 1] + 1e-6
    by0, byN = b[:, 2] - 1e-6, b[:, 3] + 1e-6
    ok = np.ones(len(qi), dtype=bool)
    for c, pt, b0, bN in ((ra[qi], ra[pj], bx0, bxN), (dec[qi], dec[pj], by0, byN)):
        i_lo = _clamp(digitize(c - theta[qi], b0, bN))
        i_hi = _clamp(digitize(c + theta[qi], b0, bN))
        ip = digitize(pt, b0, bN)
        ok &= (ip >= i_lo) & (ip <= i_hi)
    N = np.bincount(qi[ok], minlength=n).astype(float)
    return N


def candidate_center_search(galaxy_arr, p=None, mode="fast"):
    """fof.py:24-95 without the CSV side effect. Returns (main_arr (n,8), candidate_centers (m,8))."""
    p = p or Params()
    galaxy_arr = np.asarray(galaxy_arr, dtype=float)
    main_arr = np.column_stack([galaxy_arr, np.zeros(galaxy_arr.shape[0])])
    if mode == "literal":
        main_arr[:, -1] = number_counts_literal(galaxy_arr, p)
    else:
        main_arr[:, -1] = number_counts_fast(galaxy_arr, p)
    cand = main_arr[main_arr[:, -1] >= p.min_count]
    cand = cand[cand[:, -1].argsort(kind="stable")[::-1]]  # pin D1
    return main_arr, cand


# ------------------------------------------------------------------------------------------------
# stages 1-4: FoF (fof.py:98-392)
# ------------------------------------------------------------------------------------------------
def _isin_rows(rows, member_matrix):
    """fof.py:217-223 (Q5): keep rows none of whose 8 values occurs anywhere in member_matrix."""
    mask = np.isin(rows, member_matrix, invert=True)
    return np.isin(mask.sum(axis=1), rows.shape[1])


def center_members(center, galaxy_data, p, widx=None, stats=None):
    """fof.py:158-228 for one centre. Returns (member rows or None, n_virial)."""
    zc = center[2]
    if widx is None:  # literal: fof.py:158-162
        mask = np.abs(redshift_to_velocity(galaxy_data[:, 2], zc)) <= p.max_velocity
        velocity_bin = galaxy_data[mask]
        bbox = bbox_of(velocity_bin[:, 0], velocity_bin[:, 1]) if len(velocity_bin) else None
    else:
        lo, hi = widx.window(zc)
        bbox = widx.bbox(lo, hi) if hi > lo else None
    # fof.py:165-175: proper virial radius -> angle -> comoving length -> angle  (= theta * (1+z), Q4)
    ang_virial_radius = np.deg2rad(float(linear_to_angular_dist(p.virial_radius, zc)))
    max_dist_mpc = ang_virial_radius * float(comoving_distance(zc))
    max_dist = float(proper_to_deg(max_dist_mpc, zc))
    if bbox is None:
        return None, 0
    if widx is None:
        pos = grid_query(velocity_bin[:, 0], velocity_bin[:, 1], bbox, center[0], center[1], max_dist)
        virial_points = velocity_bin[pos]
    else:
        cand = widx.candidates(center[0], center[1], max_dist, lo, hi)
        pos = grid_query(galaxy_data[cand, 0], galaxy_data[cand, 1], bbox, center[0], center[1], max_dist)
        virial_points = galaxy_data[cand[pos]]
    n_v = len(virial_points)
    if n_v < p.min_virial:
        return None, n_v
    mean_sep = mean_separation(n_v, zc, p.max_velocity, p.survey_area)
    linking_length = float(proper_to_deg(p.linking_length_factor * mean_sep, zc))
    vra, vdec = virial_points[:, 0], virial_points[:, 1]
    vbox = bbox_of(vra, vdec)
    f_pos = grid_query(vra, vdec, vbox, center[0], center[1], linking_length)
    f_points = virial_points[f_pos]
    member_galaxies = f_points
    oof = bool(np.all(out_of_field(vbox, f_points[:, 0], f_points[:, 1], linking_length))) if len(f_points) else True
    for f in f_points:  # fof.py:213-228
        idx = grid_query(vra, vdec, vbox, f[0], f[1], linking_length, out_of_field_all=oof)
        fof_points = virial_points[idx]
        if len(fof_points) == 0:
            continue
        keep = _isin_rows(fof_points, member_galaxies)
        fof_points = fof_points[keep].reshape((-1, center.shape[0]))
        if len(fof_points):
            if stats is not None:
                stats["hop2_admitted"] = stats.get("hop2_admitted", 0) + len(fof_points)
            member_galaxies = np.concatenate((member_galaxies, fof_points))
    return member_galaxies, n_v


def find_candidates(galaxy_data, candidate_centers, p, widx=None, stats=None):
    """fof.py:145-246: greedy, tracker-gated loop over centres in priority order."""
    candidates = []
    tracker = np.ones(len(candidate_centers))
    luminous_gal_id = candidate_centers[:, 4]  # Q3: column 4
    for i, center in enumerate(candidate_centers):
        if not tracker[i]:
            continue
        members, _ = center_members(center, galaxy_data, p, widx, stats)
        if members is None or len(members) < p.richness:
            continue
        candidates.append(Cluster(center, members))
        _, _, coi_idx = np.intersect1d(members[:, 4], luminous_gal_id, return_indices=True)
        tracker[coi_idx] = 0
    return candidates


def setdiff2d(cluster1, cluster2):
    """helper.py:138-146."""
    return cluster1[np.isin(cluster1[:, 6], cluster2[:, 6], invert=True), :]


def overlap_removal(candidates, p, stats=None):
    """fof.py:258-314.  Contests mutate the shared Cluster objects in place (Q6)."""
    if not candidates:
        return []
    cc = np.array([[c.ra, c.dec, c.z, c.gal_id] for c in candidates])
    gal_id_space = cc[:, 3]
    first_by_id = {}
    for k, g in enumerate(gal_id_space.tolist()):
        first_by_id.setdefault(g, k)
    for center in candidates:
        mask = np.abs(redshift_to_velocity(cc[:, 2], center.z)) <= p.max_velocity
        vb = cc[mask]
        max_dist = float(linear_to_angular_dist(p.virial_radius, center.z))
        pos = grid_query(vb[:, 0], vb[:, 1], bbox_of(vb[:, 0], vb[:, 1]), center.ra, center.dec, max_dist)
        for row in vb[pos]:
            c = candidates[first_by_id[row[-1]]]
            if center.gal_id == c.gal_id:
                continue
            if len(c.galaxies) and len(center.galaxies):
                if len(setdiff2d(c.galaxies, center.galaxies)):
                    if stats is not None:
                        stats["contests"] = stats.get("contests", 0) + 1
                    c_gal, center_gal = c.remove_overlap(center)
                    c.galaxies = c_gal
                    center.galaxies = center_gal
    return [c for c in candidates if c.richness >= p.richness]


def bcg_recentre(merged, p, stats=None):
    """fof.py:318-371, literally (practically dead under the pipeline column layout, Q7)."""
    bcg_clusters = list(merged)
    for center in sorted(merged, key=lambda x: x.N, reverse=True):
        space = np.array([c.gal_id for c in bcg_clusters])
        g = center.galaxies
        max_dist = p.bcg_fraction * float(linear_to_angular_dist(p.virial_radius, center.z))
        pos = grid_query(g[:, 0], g[:, 1], bbox_of(g[:, 0], g[:, 1]), center.ra, center.dec, max_dist)
        bcg_arr = g[pos]
        if len(bcg_arr) and len(g):
            mag_sort = bcg_arr[bcg_arr[:, 3].argsort()]
            mag_sort = mag_sort[np.isin(mag_sort[:, 4], space, invert=True)]
            if len(mag_sort):
                bcg = mag_sort[0]
                if (abs(bcg[3]) > abs(center.bcg_absMag)) and (bcg[4] != center.gal_id):
                    if stats is not None:
                        stats["bcg_fired"] = stats.get("bcg_fired", 0) + 1
                    # fof.py:361-364: np.where(list == scalar) is empty => nothing is deleted
                    bcg_clusters.append(Cluster(bcg, center.galaxies))
    return [c for c in bcg_clusters if c.richness >= p.richness]


def find_number_count(center_ra, center_dec, center_z, ra, dec, distance, bbox=None):
    """helper.py:149-188."""
    theta = float(linear_to_angular_dist(distance, center_z))
    bbox = bbox or bbox_of(ra, dec)
    return len(grid_query(ra, dec, bbox, center_ra, center_dec, theta))


def center_overdensity(center, galaxy_data, p, widx=None):
    """helper.py:191-221; the 2x300 uniforms come from numpy's legacy global RNG (pin D3)."""
    n = p.n_random
    ra_random = np.random.uniform(low=min(galaxy_data[:, 0]), high=max(galaxy_data[:, 0]), size=n)
    dec_random = np.random.uniform(low=min(galaxy_data[:, 1]), high=max(galaxy_data[:, 1]), size=n)
    theta = float(linear_to_angular_dist(p.count_radius, center.z))
    if widx is None:
        mask = np.abs(redshift_to_velocity(galaxy_data[:, 2], center.z)) <= p.max_velocity
        vb = galaxy_data[mask]
        bbox = bbox_of(vb[:, 0], vb[:, 1])
        oof = bool(np.all(out_of_field(bbox, ra_random, dec_random, theta)))
        N_arr = np.array([len(grid_query(vb[:, 0], vb[:, 1], bbox, a, d, theta, out_of_field_all=oof))
                          for a, d in zip(ra_random, dec_random)])
    else:
        lo, hi = widx.window(center.z)
        bbox = widx.bbox(lo, hi)
        oof = bool(np.all(out_of_field(bbox, ra_random, dec_random, theta)))
        N_arr = np.zeros(n, dtype=np.int64)
        for k, (a, d) in enumerate(zip(ra_random, dec_random)):
            cand = widx.candidates(a, d, theta, lo, hi)
            N_arr[k] = len(grid_query(galaxy_data[cand, 0], galaxy_data[cand, 1], bbox, a, d, theta,
                                      out_of_field_all=oof))
    N_mean = np.mean(N_arr)
    N_rms = np.sqrt(np.mean(N_arr ** 2))
    return (center.N - N_mean) / N_rms


def select_clusters(bcg_clusters, galaxy_data, p, widx=None):
    """fof.py:374-392."""
    final = []
    for center in bcg_clusters:
        g = center.galaxies
        center.N = find_number_count(center.ra, center.dec, center.z, g[:, 0], g[:, 1], p.count_radius)
        center.D = center_overdensity(center, galaxy_data, p, widx)
        if (center.N >= p.min_count) and (center.richness >= p.richness) and (center.D >= p.overdensity):
            final.append(center)
    return final


def FoF(galaxy_data, candidate_centers, p=None, mode="fast", stats=None, stages=None):
    """fof.py:98-392.  ``stages`` (dict) receives the intermediate cluster lists for parity tests."""
    p = p or Params()
    galaxy_data = np.asarray(galaxy_data, dtype=float)
    candidate_centers = np.asarray(candidate_centers, dtype=float)
    widx = WindowIndex(galaxy_data, p.max_velocity) if mode == "fast" else None
    candidates = find_candidates(galaxy_data, candidate_centers, p, widx, stats)
    if stages is not None:
        stages["candidates"] = [(c.center_attributes, c.galaxies.copy()) for c in candidates]
    merged = overlap_removal(candidates, p, stats)
    if stages is not None:
        stages["merged"] = [c.gal_id for c in merged]
    bcg = bcg_recentre(merged, p, stats)
    return select_clusters(bcg, galaxy_data, p, widx)


# ------------------------------------------------------------------------------------------------
# stage 5: interloper removal (fof.py:396-516)
# ------------------------------------------------------------------------------------------------
def three_sigma(center, member_galaxies, n):
    """fof.py:396-496."""
    N = member_galaxies.shape
    mg = np.column_stack([member_galaxies, np.zeros((N[0], 2))])
    mg[:, -2] = redshift_to_velocity(mg[:, 2], center.z)
    sep = vincenty_rad(np.deg2rad(center.ra), np.deg2rad(center.dec), np.deg2rad(mg[:, 0]), np.deg2rad(mg[:, 1]))
    d_A = float(angular_diameter_distance(center.z))
    mg[:, -1] = np.rad2deg(sep) * d_A * (math.pi / 180.0)
    cleaned = []
    s = mg[:, -1]
    bins = np.linspace(min(s), max(s), n)
    digitized = np.digitize(s, bins)
    for i in range(1, len(bins) + 1):
        b = mg[np.where(digitized == i)]
        size_before = len(b)
        size_change = 0 - size_before
        while (size_change != 0) and (len(b) > 0):
            size_before = len(b)
            if len(b) > 2:
                vel = b[:, -2]
                # pin D4: exact ties in |v - mean| (duplicated redshifts) keep their order; the reference's default
                # argsort leaves them to the sort implementation (measure zero for real data)
                b = b[np.argsort(np.abs(vel - np.mean(vel)), kind="stable")]
                if b[-1, -1] == 0.0:
                    ignored = b[-2, :]
                    new_b = np.delete(b, (-2), axis=0)
                else:
                    ignored = b[-1, :]
                    new_b = b[:-1, :]
                new_vel = new_b[:, -2]
                if b[-1, -1] != 0.0:
                    if len(new_vel) > 1:
                        bin_mean = np.mean(new_vel)
                        bin_disp = sum((new_vel - bin_mean) ** 2) / (len(new_vel) - 1)
                        if abs(ignored[-2] - bin_mean) > 3 * np.sqrt(bin_disp):
                            b = new_b
            size_change = len(b) - size_before
        cleaned.append(b[:, :-2])
    return np.concatenate(cleaned, axis=0)


def interloper_removal(data, p=None):
    """fof.py:499-516 (mutates c.galaxies like the reference)."""
    p = p or Params()
    removed = []
    for c in data:
        initial = len(c.galaxies)
        c.galaxies = three_sigma(c, c.galaxies, p.n_bins)
        removed.append(initial - len(c.galaxies))
    return [c for c in data if c.richness >= p.richness], removed


# ------------------------------------------------------------------------------------------------
# stage 6: virial estimator (analysis/mass_estimator.py:36-112,148-184; mass_analysis.py:8-44)
# ------------------------------------------------------------------------------------------------
def velocity_error(galaxies, z_arr, z_average, velocities):
    """analysis/mass_estimator.py:148-167."""
    zerr = np.amax(abs(galaxies[:, 3:5] - z_arr[:, np.newaxis]), axis=1)
    z_average_err = np.sqrt(sum(zerr ** 2)) / len(z_arr)
    z_diff_err = np.sqrt(np.add(z_average_err ** 2, zerr ** 2))
    return abs(velocities) * np.sqrt((z_average_err / z_average) ** 2 + (z_diff_err / (z_arr - z_average)) ** 2)


def bootstrap_dispersions(vel):
    """analysis/mass_estimator.py:178-184 + astropy bootstrap / NumpyRNGContext(1)."""
    state = np.random.get_state()
    np.random.seed(1)
    try:
        n = len(vel)
        out = np.empty(100)
        for i in range(100):
            x = vel[np.random.randint(low=0, high=n, size=n - 1)]
            out[i] = np.sum(x ** 2) / (len(x) - 1)
    finally:
        np.random.set_state(state)
    return out


def virial_mass_estimator(galaxies):
    """analysis/mass_estimator.py:36-112. Returns (M [Msun/h], M_err, sigma^2 [km^2/s^2], err, R [Mpc/h])."""
    n = len(galaxies)
    z_arr = galaxies[:, 2]
    zbar = np.mean(z_arr)
    vel = redshift_to_velocity(z_arr, zbar)
    vel_err = velocity_error(galaxies, z_arr, zbar, vel)
    vel_disp = np.sum(vel ** 2) / (n - 1)
    bootstrap_disp = np.std(bootstrap_dispersions(vel))
    harmonic_disp = (1 / n) * (sum(1 / (vel_err ** 2))) ** (-1)
    combined = np.sqrt(harmonic_disp ** 2 + bootstrap_disp ** 2)
    lon = np.deg2rad(galaxies[:, 0])
    lat = np.deg2rad(galaxies[:, 1])
    iu, ju = np.triu_indices(n, k=1)  # itertools.combinations order
    sep_deg = np.rad2deg(vincenty_rad(lon[iu], lat[iu], lon[ju], lat[ju]))
    d_A = float(angular_diameter_distance(zbar))
    actual = sep_deg * d_A * (math.pi / 180.0) * LITTLE_H
    total = 0.0
    for x in (1 / actual).tolist():  # python sum(): sequential, left to right
        total += x
    radius = n * (n - 1) / total
    mass = (3 / 2) * np.pi * (vel_disp * radius) * MASS_UNIT
    return mass, mass * (combined / vel_disp), vel_disp, combined, radius


def projected_mass_estimator(galaxies):
    """analysis/mass_estimator.py:115-145. Returns (M [Msun/h], M_err [km^2/s^2])."""
    N = len(galaxies)
    z_arr = galaxies[:, 2]
    zbar = np.mean(z_arr)
    vel = redshift_to_velocity(z_arr, zbar)
    lon, lat = np.deg2rad(galaxies[:, 0]), np.deg2rad(galaxies[:, 1])
    clon, clat = np.deg2rad(np.mean(galaxies[:, 0])), np.deg2rad(np.mean(galaxies[:, 1]))
    sep_deg = np.rad2deg(vincenty_rad(clon, clat, lon, lat))
    d_A = float(angular_diameter_distance(zbar))
    actual = sep_deg * d_A * (math.pi / 180.0) * LITTLE_H
    sumproduct = np.sum(actual * vel ** 2)
    mass = (32 / np.pi) * (1 / (N - 1.5)) * sumproduct * MASS_UNIT
    vel_err = velocity_error(galaxies, z_arr, zbar, vel)
    mass_err = np.sqrt(sum((2 * vel * vel_err) ** 2))
    return float(mass), float(mass_err)


def find_masses(data):
    """mass_analysis.py:8-44, estimator == 'virial'. Returns a list of dicts."""
    out = []
    for c in data:
        mass, mass_err, vd, vd_err, radius = virial_mass_estimator(c.galaxies)
        absmag = c.galaxies[:, 5]
        total_absmag = -2.5 * np.log10(sum(10 ** (-0.4 * absmag)))
        out.append(dict(center=c.center_attributes, galaxies=c.galaxies, cluster_mass=float(mass),
                        mass_err=float(mass_err), vel_disp=float(vd), vel_disp_err=float(vd_err),
                        projected_radius=float(radius), total_absmag=float(total_absmag),
                        D=float(c.D), N=float(c.N)))
    return out


def run_pipeline(df_values, p=None, mode="fast", seed=12345, zfilter=(0.5, 2.5), stats=None):
    """pipeline.py:62-64,80-99,113,173 call order; same artefact names as oracle/run_reference.py."""
    p = p or Params()
    out = {}
    main_arr, centers = candidate_center_search(df_values, p, mode)
    out["main_arr"], out["candidate_centers"] = main_arr, centers
    g = main_arr[(main_arr[:, 2] >= zfilter[0]) & (main_arr[:, 2] <= zfilter[1] + 0.02)]
    c = centers[(centers[:, 2] >= zfilter[0]) & (centers[:, 2] <= zfilter[1])]
    np.random.seed(seed)
    clusters = FoF(g, c, p, mode, stats)
    out["clusters"] = [(cl.center_attributes.copy(), cl.galaxies.copy(), float(cl.N), float(cl.D)) for cl in clusters]
    cleaned, removed = interloper_removal(clusters, p)
    out["cleaned"] = [(cl.center_attributes.copy(), cl.galaxies.copy()) for cl in cleaned]
    out["number_removed"] = removed
    out["virial"] = find_masses(cleaned)
    return out


# ------------------------------------------------------------------------------------------------
# Catalogue matching (catalog_matching.py:21-111) -- SURVEY section 8(f) item 3
# ------------------------------------------------------------------------------------------------
def matching_angle(d_mpc_h, z):
    """catalog_matching.py:85-91: proper length -> angle -> comoving length (x D_M, flat: D_M = D_C) -> angle again;
    degrees."""
    first = linear_to_angular_dist(d_mpc_h, z)
    comoving_mpc = np.deg2rad(first) * comoving_distance(z)
    return proper_to_deg(comoving_mpc, z)


def unit_vectors(ra_deg, dec_deg):
    """astropy UnitSphericalRepresentation.to_cartesian."""
    lon, lat = np.deg2rad(ra_deg), np.deg2rad(dec_deg)
    return np.stack([np.cos(lat) * np.cos(lon), np.cos(lat) * np.sin(lon), np.sin(lat)], axis=-1)


def compare_clusters(sample, catalog, matching_distance, z_cut=0.1, detail=False):
    """catalog_matching.py:21-111 without logging / plotting.

    sample: (S, 4) rows [ra, dec, z, richness]; catalog: (C, >=4) rows [ra, dec, z, richness, ...];
    matching_distance in h^-1 Mpc.  Returns (sample_matched, catalog_matching_idx) over the catalogue rows that pass
    the redshift filter of :66-68; detail=True adds (kept catalogue rows, matched sample index or -1, separation in
    degrees of the nearest sample row or nan).

    Pinned order D5: the nearest sample row is the one with the smallest squared chord between unit vectors,
    ((dx*dx + dy*dy) + dz*dz) in float64, lowest index on exact ties (astropy asks a scipy cKDTree, whose tie order
    is unspecified)."""
    sample = np.asarray(sample, dtype=np.float64)
    catalog = np.asarray(catalog)
    z_arr = sample[:, 2]
    lo, hi = min(z_arr), max(z_arr)
    sample = sample[(sample[:, 2] >= lo - 0.1) & (sample[:, 2] <= hi + 0.1)]
    kept = np.flatnonzero((catalog[:, 2] >= lo - 0.1) & (catalog[:, 2] <= hi + 0.1))
    catalog = catalog[kept]
    shape = len(catalog)
    sample_matched = np.zeros((shape, 4))
    catalog_matching_idx = np.zeros(shape)
    match_row = np.full(shape, -1, dtype=np.int64)
    match_sep = np.full(shape, np.nan)
    sxyz = unit_vectors(sample[:, 0], sample[:, 1])
    for i, c in enumerate(catalog):
        cz = float(c[2])
        rows = np.flatnonzero(np.abs(sample[:, 2] - cz) <= z_cut)
        if not len(rows):
            continue
        max_sep = matching_angle(matching_distance, cz)
        cxyz = unit_vectors(float(c[0]), float(c[1]))
        d = sxyz[rows] - cxyz
        chord2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        k = rows[int(np.argmin(chord2))]
        sep = np.rad2deg(vincenty_rad(np.deg2rad(sample[k, 0]), np.deg2rad(sample[k, 1]),
                                      np.deg2rad(float(c[0])), np.deg2rad(float(c[1]))))
        match_sep[i] = sep
        if sep <= max_sep:
            catalog_matching_idx[i] = 1
            sample_matched[i, :] = sample[k, :]
            match_row[i] = k
    sample_matched = sample_matched[np.all(sample_matched != 0, axis=1)]
    if detail:
        return sample_matched, catalog_matching_idx, kept, match_row, match_sep
    return sample_matched, catalog_matching_idx
