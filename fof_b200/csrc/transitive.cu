// Transitive friends-of-friends (connected components) -- kernel K14 of DESIGN.md.
//
// NOT a reference code path: kencx/FoF never computes connected components.  This is
// the lock-free union-find named in BASELINE.json's north_star, offered as a separately named mode with its own
// oracle (oracle/transitive_oracle.py, scipy connected components on the identical predicate).
//
// Link predicate (symmetric): galaxies i and j are friends iff each lies in the other's velocity window
// (|v(z_j; z_i)| <= v_max and |v(z_i; z_j)| <= v_max, analysis/methods.py:68-87) and their haversine separation is
// within the angle subtended by the transverse linking length at BOTH redshifts
// (d <= theta(ll, z_i) and d <= theta(ll, z_j), analysis/methods.py:12-32).  Groups are the connected components;
// the label of a galaxy is the smallest row index in its component.
#include <cub/cub.cuh>

#include "catalog_host.cuh"
#include "stage0.cuh"
#include "transitive.cuh"

namespace fofb {

__global__ void k_link_prep(CatalogView cat, Cosmo cosmo, double ll_mpc_h, double vmax, QueryRec* rec, double* theta_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= cat.n) return;
  const double z = cat.zs[r];
  const double theta = cosmo_linear_to_angular_dist(cosmo, ll_mpc_h, z);
  int lo, hi;
  velocity_window(cat.zs, (int)cat.n, z, vmax, &lo, &hi);
  QueryRec q;
  q.r_deg = theta;
  q.wlo = lo;
  q.whi = hi;
  q.ra_lo = q.dec_lo = -INFINITY;
  q.ra_hi = q.dec_hi = INFINITY;
  rec[r] = q;
  theta_out[r] = theta;
}

__global__ void k_uf_init(int n, int* parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) parent[i] = i;
}

// find with path halving; parent[] only ever decreases, so stale reads are safe
__device__ __forceinline__ int uf_find(int* parent, int x) {
  for (;;) {
    const int p = ((volatile int*)parent)[x];
    if (p == x) return x;
    const int gp = ((volatile int*)parent)[p];
    if (gp != p) atomicMin(parent + x, gp);
    x = p;
  }
}

// lock-free union: hook the larger root under the smaller one with atomicCAS, retry on interference
__device__ __forceinline__ void uf_union(int* parent, int a, int b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    if (atomicCAS(parent + b, b, a) == b) return;
  }
}

// One thread per galaxy in cell order; every accepted pair is seen from both ends, the end with the larger row
// index performs the union.
template <bool COUNT_LINKS>
__global__ void __launch_bounds__(256) k_link_union(CellListView cl, int n, const QueryRec* __restrict__ rec, int* parent,
                                                    unsigned long long* n_links) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int zr = cl.zrank[q];
  const QueryRec qr = rec[zr];
  const double ra = cl.ra[q], dec = cl.dec[q], cs = cl.cosdec[q];
  const int row = cl.row[q];
  BubbleTest bt;
  bt.init(ra, dec, cs, qr.r_deg);
  const int cx = cl.cx_of(ra), cy = cl.cy_of(dec);
  const int x0 = cx > 0 ? cx - 1 : 0, x1 = cx < cl.ncx - 1 ? cx + 1 : cl.ncx - 1;
  const int y0 = cy > 0 ? cy - 1 : 0, y1 = cy < cl.ncy - 1 ? cy + 1 : cl.ncy - 1;
  unsigned links = 0;
  for (int yy = y0; yy <= y1; ++yy) {
    int s = cl.cell_start[yy * cl.ncx + x0];
    for (int xx = x0; xx <= x1; ++xx) {
      const int e = cl.cell_start[yy * cl.ncx + xx + 1];
      int a = s, b = e;
      while (a < b) { const int mid = (a + b) >> 1; if (cl.zrank[mid] < qr.wlo) a = mid + 1; else b = mid; }
      const int first = a;
      b = e;
      while (a < b) { const int mid = (a + b) >> 1; if (cl.zrank[mid] < qr.whi) a = mid + 1; else b = mid; }
      for (int j = first; j < a; ++j) {
        const int rj = cl.row[j];
        if (rj >= row) continue;
        if (!COUNT_LINKS && uf_find(parent, row) == uf_find(parent, rj)) continue;
        const int zj = cl.zrank[j];
        const QueryRec other = rec[zj];
        if (zr < other.wlo || zr >= other.whi) continue;       // i must lie in j's window too
        const double pra = cl.ra[j], pdec = cl.dec[j], pcs = cl.cosdec[j];
        if (!bt.inside(pra, pdec, pcs)) continue;              // d <= theta_i
        BubbleTest bo;
        bo.init(pra, pdec, pcs, other.r_deg);
        if (!bo.inside(ra, dec, cs)) continue;                 // d <= theta_j
        uf_union(parent, row, rj);
        ++links;
      }
      s = e;
    }
  }
  if (links) atomicAdd(n_links, (unsigned long long)links);
}

// ---- run-based linking (no link count) --------------------------------------------------------------------------
// With cells whose diagonal is below the smallest linking angle, the members of one cell that also share a bin of
// ln(1 + z) of width ln(1 + v_max / c) are friends of each other by construction: every pair is closer than both
// linking angles and inside both velocity windows.  Such a run (contiguous in the cell list, which is ordered by
// redshift rank inside a cell) is united with its first member without testing pairs, and a galaxy then needs ONE
// link into a run -- or one look at the run's root -- instead of a test against each member.  A percolating blob of
// 10^5 galaxies costs O(runs) per galaxy instead of O(members).
__global__ void k_run_starts(CellListView cl, CatalogView cat, int n, double inv_w, int* start_pos) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  bool start = q == 0;
  if (!start) {
    const int c1 = cl.cy_of(cl.dec[q]) * cl.ncx + cl.cx_of(cl.ra[q]);
    const int c0 = cl.cy_of(cl.dec[q - 1]) * cl.ncx + cl.cx_of(cl.ra[q - 1]);
    const double u1 = floor(log1p(cat.zs[cl.zrank[q]]) * inv_w), u0 = floor(log1p(cat.zs[cl.zrank[q - 1]]) * inv_w);
    start = c1 != c0 || !(u1 == u0);
  }
  start_pos[q] = start ? q : 0;
}

__global__ void k_run_ends(int n, const int* head, int* run_end) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  if (q == n - 1) run_end[head[q]] = n;
  if (q > 0 && head[q] == q) run_end[head[q - 1]] = q;
}

__global__ void k_run_union(CellListView cl, int n, const int* head, int* parent) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int hq = head[q];
  if (hq != q) uf_union(parent, cl.row[q], cl.row[hq]);
}

struct MaxOp {
  __host__ __device__ int operator()(int a, int b) const { return a > b ? a : b; }
};

// One thread per galaxy in cell order over the cells its bubble can reach.  FINE: runs are already united, so a run
// whose root equals the galaxy's is skipped as a whole and the first link into a run ends the visit of that run.
// Every friend pair is seen from both ends and tested from the end with the larger row index, as in k_link_union.
template <bool FINE>
__global__ void __launch_bounds__(256) k_link_union_runs(CellListView cl, int n, const QueryRec* __restrict__ rec,
                                                         const int* __restrict__ head, const int* __restrict__ run_end,
                                                         int* parent) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int zr = cl.zrank[q];
  const QueryRec qr = rec[zr];
  const double ra = cl.ra[q], dec = cl.dec[q], cs = cl.cosdec[q];
  const int row = cl.row[q];
  BubbleTest bt;
  bt.init(ra, dec, cs, qr.r_deg);
  const double reach = ra_reach(qr.r_deg, dec);
  const int x0 = cl.cx_of(ra - reach), x1 = cl.cx_of(ra + reach);
  const int y0 = cl.cy_of(dec - qr.r_deg), y1 = cl.cy_of(dec + qr.r_deg);
  for (int yy = y0; yy <= y1; ++yy) {
    for (int xx = x0; xx <= x1; ++xx) {
      const int s = cl.cell_start[yy * cl.ncx + xx], e = cl.cell_start[yy * cl.ncx + xx + 1];
      int a = s, b = e;
      while (a < b) { const int mid = (a + b) >> 1; if (cl.zrank[mid] < qr.wlo) a = mid + 1; else b = mid; }
      int j = a;
      b = e;
      while (a < b) { const int mid = (a + b) >> 1; if (cl.zrank[mid] < qr.whi) a = mid + 1; else b = mid; }
      const int last = a;
      while (j < last) {
        int stop = last;
        if (FINE) {
          const int hj = head[j];
          const int re = run_end[hj];
          stop = re < last ? re : last;
          if (uf_find(parent, row) == uf_find(parent, cl.row[hj])) { j = stop; continue; }
        }
        for (; j < stop; ++j) {
          const int rj = cl.row[j];
          if (rj >= row) continue;
          if (!FINE && uf_find(parent, row) == uf_find(parent, rj)) continue;
          const int zj = cl.zrank[j];
          const QueryRec other = rec[zj];
          if (zr < other.wlo || zr >= other.whi) continue;       // i must lie in j's window too
          const double pra = cl.ra[j], pdec = cl.dec[j], pcs = cl.cosdec[j];
          if (!bt.inside(pra, pdec, pcs)) continue;              // d <= theta_i
          BubbleTest bo;
          bo.init(pra, pdec, pcs, other.r_deg);
          if (!bo.inside(ra, dec, cs)) continue;                 // d <= theta_j
          uf_union(parent, row, rj);
          if (FINE) break;                                       // the rest of the run hangs on rj already
        }
        if (FINE) j = stop;
      }
    }
  }
}

// pointer jumping to the root = smallest row index of the component
__global__ void k_uf_labels(int n, int* parent, int* label, int* is_root) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x = i;
  while (parent[x] != x) x = parent[x];
  label[i] = x;
  is_root[i] = x == i ? 1 : 0;
}

void Transitive::run(fofb_handle* h, const fofb_params& p, const View& rows, double ll_mpc_h, bool count_links) {
  cudaStream_t st = h->stream;
  FOFB_REQUIRE(ll_mpc_h > 0 && std::isfinite(ll_mpc_h), "linking length must be positive");
  cat.build(h, rows);
  const int n = (int)cat.n;
  const int B = 256;
  const Cosmo cosmo = device_cosmo(h, p);
  QueryRec* recs = rec.alloc(n);
  double* th = theta.alloc(n);
  FOFB_PROFILED(h, "k_link_prep", k_link_prep<<<grid_for(n, B), B, 0, st>>>(cat.view(), cosmo, ll_mpc_h, p.max_velocity, recs, th));
  double* tmax = theta_max.alloc(1);
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceReduce::Max(nullptr, bytes, th, tmax, n, st));
  FOFB_CUDA(cub::DeviceReduce::Max(h->cub_tmp.alloc(bytes + 256), bytes, th, tmax, n, st));
  count_launch(h, 2);
  double r_max = 0.0;
  FOFB_CUDA(cudaMemcpyAsync(&r_max, tmax, sizeof(double), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  FOFB_REQUIRE(std::isfinite(r_max) && r_max > 0.0, "non-finite linking angle");
  int* par = parent.alloc(n);
  unsigned long long* nl = links.alloc(1);
  FOFB_CUDA(cudaMemsetAsync(nl, 0, sizeof(unsigned long long), st));
  k_uf_init<<<grid_for(n, B), B, 0, st>>>(n, par);
  FOFB_LAUNCH_CHECK(h);
  if (count_links) {
    cells.build(h, cat, r_max);
    FOFB_PROFILED(h, "k_link_union", k_link_union<true><<<grid_for(n, 256), 256, 0, st>>>(cells.view(), n, recs, par, nl));
  } else {
    // cells with a diagonal below the smallest linking angle (the table-size cap of CellList::build may enlarge them)
    FOFB_CUDA(cub::DeviceReduce::Min(nullptr, bytes, th, tmax, n, st));
    FOFB_CUDA(cub::DeviceReduce::Min(h->cub_tmp.alloc(bytes + 256), bytes, th, tmax, n, st));
    count_launch(h, 2);
    double r_min = 0.0;
    FOFB_CUDA(cudaMemcpyAsync(&r_min, tmax, sizeof(double), cudaMemcpyDeviceToHost, st));
    FOFB_CUDA(cudaStreamSynchronize(st));
    FOFB_REQUIRE(std::isfinite(r_min) && r_min > 0.0, "non-finite linking angle");
    const double stretch = 1.01 / cos(fmin(89.0, cat.dec_absmax + r_min) * kDeg2Rad);
    cells.build(h, cat, 0.94 * r_min / sqrt(1.0 + stretch * stretch));
    const bool fine = p.max_velocity > 0.0 && hypot(cells.s_ra, cells.s_dec) <= 0.95 * r_min;
    int* hd = run_head.alloc(n);
    int* re = run_end.alloc(n);
    if (fine) {
      int* sp = run_start.alloc(n);
      const double w = log1p(p.max_velocity / kCkms) * (1.0 - 1.0e-6);
      k_run_starts<<<grid_for(n, B), B, 0, st>>>(cells.view(), cat.view(), n, 1.0 / w, sp);
      FOFB_LAUNCH_CHECK(h);
      FOFB_CUDA(cub::DeviceScan::InclusiveScan(nullptr, bytes, sp, hd, MaxOp(), n, st));
      FOFB_CUDA(cub::DeviceScan::InclusiveScan(h->cub_tmp.alloc(bytes + 256), bytes, sp, hd, MaxOp(), n, st));
      count_launch(h, 2);
      k_run_ends<<<grid_for(n, B), B, 0, st>>>(n, hd, re);
      FOFB_LAUNCH_CHECK(h);
      FOFB_PROFILED(h, "k_run_union", k_run_union<<<grid_for(n, B), B, 0, st>>>(cells.view(), n, hd, par));
      FOFB_PROFILED(h, "k_link_union", k_link_union_runs<true><<<grid_for(n, 256), 256, 0, st>>>(cells.view(), n, recs, hd, re, par));
    } else {
      FOFB_PROFILED(h, "k_link_union", k_link_union_runs<false><<<grid_for(n, 256), 256, 0, st>>>(cells.view(), n, recs, hd, re, par));
    }
    used_runs = fine;
  }
  int* lab = label.alloc(n);
  int* roots = is_root.alloc(n);
  FOFB_PROFILED(h, "k_uf_labels", k_uf_labels<<<grid_for(n, B), B, 0, st>>>(n, par, lab, roots));
  int* d_groups = num_groups.alloc(1);
  FOFB_CUDA(cub::DeviceReduce::Sum(nullptr, bytes, roots, d_groups, n, st));
  FOFB_CUDA(cub::DeviceReduce::Sum(h->cub_tmp.alloc(bytes + 256), bytes, roots, d_groups, n, st));
  count_launch(h, 2);
  int g = 0;
  unsigned long long l = 0;
  FOFB_CUDA(cudaMemcpyAsync(&g, d_groups, sizeof(int), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaMemcpyAsync(&l, nl, sizeof(l), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  n_groups = g;
  n_links = count_links ? (long long)l : -1;
  n_rows = n;
}

}  // namespace fofb
