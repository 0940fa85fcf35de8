"""Stand-in for GriSPy (mchalela/GriSPy 0.2.x) -- oracle test infrastructure only.

GriSPy is an un-vendored, un-pinned dependency of the reference (README.md:28) and is not
installed here, so this module RESTATES its published algorithm for the one configuration the
reference uses: ``GriSPy(data, metric="haversine")`` followed by
``bubble_neighbors(centres, distance_upper_bound=r)``  (call sites fof.py:162,177,202-211,275-282,
330-339; helper.py:181-185,211-213).  Parity against the real package is UNPINNED: the reference
holds no tests or golden vectors.

Restated behaviour:
 * grid: per axis ``linspace(min - 1e-6, max + 1e-6, N_cells + 1)``, N_cells = 64;
 * cell index of x: ``int(N_cells * (x - bins[0]) / (bins[-1] - bins[0]))`` (C truncation);
 * cell contents in ascending data index (pin D2: GriSPy's argsort branch for > N_cells**2
   points is an unstable sort; the oracle pins it to the stable order);
 * query: cell box ``digitize(c - r) .. digitize(c + r)`` per axis, clamped to [0, N_cells-1];
   ``np.meshgrid`` default indexing => axis-0 (RA) cell index varies fastest; cells whose
   centre is farther than r + half the cell diagonal are dropped; points kept iff
   ``haversine_deg(c, p) <= r``; results unsorted, in cell-then-index order.
"""
import numpy as np

__version__ = "0.2-shim"


def haversine(c0, points):
    """Great-circle separation in DEGREES between (lon, lat) in degrees (grispy/distances.py)."""
    lon1 = np.deg2rad(c0[0])
    lat1 = np.deg2rad(c0[1])
    lon2 = np.deg2rad(points[:, 0])
    lat2 = np.deg2rad(points[:, 1])
    sdlon = np.sin((lon2 - lon1) / 2.0)
    sdlat = np.sin((lat2 - lat1) / 2.0)
    clat1 = np.cos(lat1)
    clat2 = np.cos(lat2)
    num1 = sdlat ** 2
    num2 = clat1 * clat2 * sdlon ** 2
    sep = 2 * np.arcsin(np.sqrt(num1 + num2))
    return np.rad2deg(sep)


class GriSPy:
    def __init__(self, data=None, N_cells=64, periodic=None, metric="euclid", copy_data=False):
        data = np.asarray(data, dtype=float)
        if data.ndim != 2 or data.shape[0] == 0:
            raise ValueError("Data array has the wrong shape. Expected shape of (n, k)")
        if metric != "haversine" or periodic:
            raise NotImplementedError("shim restates only metric='haversine', non-periodic")
        self.data = data.copy() if copy_data else data
        self.N_cells = int(N_cells)
        self.metric = metric
        self.dim = data.shape[1]
        self._build_grid()

    # ---- grid ----------------------------------------------------------------------------
    def _digitize(self, x, k):
        b0 = self.k_bins[0, k]
        bN = self.k_bins[-1, k]
        return (self.N_cells * (np.asarray(x, dtype=float) - b0) / (bN - b0)).astype(np.int64)

    def _build_grid(self, epsilon=1.0e-6):
        n = len(self.data)
        self.k_bins = np.zeros((self.N_cells + 1, self.dim))
        k_digit = np.zeros(self.data.shape, dtype=np.int64)
        for k in range(self.dim):
            k_data = self.data[:, k]
            self.k_bins[:, k] = np.linspace(k_data.min() - epsilon, k_data.max() + epsilon, self.N_cells + 1)
            k_digit[:, k] = self._digitize(k_data, k)
        self.k_digit = k_digit
        # cell -> ascending data indices (dict-append branch; argsort branch pinned stable)
        compact = k_digit[:, 0] + self.N_cells * k_digit[:, 1]
        order = np.argsort(compact, kind="stable")
        sorted_compact = compact[order]
        cells, starts = np.unique(sorted_compact, return_index=True)
        ends = np.append(starts[1:], n)
        self.grid = {}
        for c, s, e in zip(cells.tolist(), starts.tolist(), ends.tolist()):
            self.grid[(c % self.N_cells, c // self.N_cells)] = order[s:e]

    # ---- queries ---------------------------------------------------------------------------
    def _neighbor_cells(self, centre, r):
        lo = np.zeros(self.dim, dtype=np.int64)
        hi = np.zeros(self.dim, dtype=np.int64)
        for k in range(self.dim):
            lo[k] = min(max(int(self._digitize(centre[k] - r, k)), 0), self.N_cells - 1)
            hi[k] = min(max(int(self._digitize(centre[k] + r, k)), 0), self.N_cells - 1)
        k_grids = np.meshgrid(*[np.arange(lo[k], hi[k] + 1) for k in range(self.dim)])
        cells = np.array([g.flatten() for g in k_grids]).T
        cell_size = self.k_bins[1, :] - self.k_bins[0, :]
        cell_radii = 0.5 * np.sum(cell_size ** 2) ** 0.5
        phys = np.array([self.k_bins[cells[:, k], k] + 0.5 * cell_size[k] for k in range(self.dim)]).T
        keep = haversine(centre, phys) < r + cell_radii
        return cells[keep]

    def bubble_neighbors(self, centres, distance_lower_bound=-1.0, distance_upper_bound=-1.0,
                         sorted=False, kind="quicksort"):
        centres = np.asarray(centres, dtype=float)
        if centres.ndim != 2 or centres.shape[1] != self.dim:
            raise ValueError("Centres array has the wrong shape. Expected shape of (n, k)")
        if sorted:
            raise NotImplementedError("shim restates only sorted=False")
        r_all = np.broadcast_to(np.asarray(distance_upper_bound, dtype=float), (len(centres),))
        # centres farther than r outside the data bbox on any axis: no neighbours if ALL are so
        out_of_field = np.zeros(len(centres), dtype=bool)
        for k in range(self.dim):
            out_of_field |= centres[:, k] - r_all > self.k_bins[-1, k]
            out_of_field |= centres[:, k] + r_all < self.k_bins[0, k]
        dists, idxs = [], []
        for centre, r in zip(centres, r_all):
            if np.all(out_of_field):
                cells = np.zeros((0, self.dim), dtype=np.int64)
            else:
                cells = self._neighbor_cells(centre, r)
            chunks = [self.grid[t] for t in map(tuple, cells.tolist()) if t in self.grid]
            ind = np.concatenate(chunks).astype(np.int64) if chunks else np.zeros(0, dtype=np.int64)
            d = haversine(centre, self.data[ind]) if len(ind) else np.zeros(0)
            m = d <= r
            dists.append(d[m])
            idxs.append(ind[m])
        return dists, idxs
