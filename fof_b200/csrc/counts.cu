// Stage 0: N(0.5) number counts for every galaxy and the candidate-centre ordering
// (kernels K5, K6 of DESIGN.md).  Replaces candidate_center_search's O(n^2) loop
// (fof.py:65-81) and find_number_count (helper.py:149-188).
#include <cub/cub.cuh>

#include "catalog_host.cuh"
#include "stage0.cuh"

namespace fofb {

// ---- K5a: per-galaxy query record, one thread per z-rank ----------------------------------------
// theta_0.5(z_i), the velocity window of galaxy i, the bounding box of that window (= the bbox of
// the GriSPy grid find_number_count builds, helper.py:181) and the resulting cell-box bounds.
__global__ void k_count_prep(CatalogView cat, Cosmo cosmo, double count_radius, double vmax, int min_window,
                             int grid_cells, QueryRec* rec, double* theta_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= cat.n) return;
  const double z = cat.zs[r];
  const double theta = cosmo_linear_to_angular_dist(cosmo, count_radius, z);
  int lo, hi;
  velocity_window(cat.zs, (int)cat.n, z, vmax, &lo, &hi);
  QueryRec q;
  q.r_deg = theta;
  q.wlo = lo;
  q.whi = hi;
  if (hi - lo >= min_window) {
    const BBox bb = range_bbox(cat, lo, hi);
    const Box b = grid_box(bb, grid_cells, cat.zra[r], cat.zdec[r], theta);
    q.ra_lo = b.ra_lo; q.ra_hi = b.ra_hi; q.dec_lo = b.dec_lo; q.dec_hi = b.dec_hi;
  } else {
    q.ra_lo = q.dec_lo = INFINITY;
    q.ra_hi = q.dec_hi = -INFINITY;
    q.whi = lo;  // empty window => N stays 0 (fof.py:71)
  }
  rec[r] = q;
  theta_out[r] = theta;
}

// ---- K5: the count.  One thread per galaxy in cell order (a warp = 32 consecutive members of one
// or two adjacent cells, so their probes and candidate loads share L1 lines).  Cells are z-rank
// sorted, so the velocity window is one contiguous sub-range per neighbour cell, found by binary
// search; only in-window candidates reach the cell-box and fp64 haversine tests.
__device__ __forceinline__ int lower_bound_i32(const int* __restrict__ a, int lo, int hi, int key) {
  while (lo < hi) {
    const int m = (lo + hi) >> 1;
    if (__ldg(a + m) < key) lo = m + 1; else hi = m;
  }
  return lo;
}

__global__ void __launch_bounds__(256) k_count(CellListView cl, int n, const QueryRec* __restrict__ rec,
                                               int* __restrict__ N_row) {
  __shared__ int run_lo[9][256], run_hi[9][256];  // per thread: the runs of its (at most 3 x 3) neighbour cells
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const QueryRec qr = rec[cl.zrank[q]];
  int count = 0;
  if (qr.whi > qr.wlo) {
    const double ra = cl.ra[q], dec = cl.dec[q];
    BubbleTest bt;
    bt.init(ra, dec, cl.cosdec[q], qr.r_deg);
    const int cx = cl.cx_of(ra), cy = cl.cy_of(dec);
    const int x0 = cx > 0 ? cx - 1 : 0, x1 = cx < cl.ncx - 1 ? cx + 1 : cl.ncx - 1;
    const int y0 = cy > 0 ? cy - 1 : 0, y1 = cy < cl.ncy - 1 ? cy + 1 : cl.ncy - 1;
    // pass 1: the in-window run of every neighbour cell (uniform work: two binary searches per cell);
    // pass 2: ONE loop over the runs back to back, so that a warp iterates max-over-lanes of the TOTAL number of
    // candidates instead of the sum over cells of the per-cell maxima (the runs are 0-3 long for field galaxies).
    int nrun = 0;
    for (int yy = y0; yy <= y1; ++yy) {
      int s = __ldg(cl.cell_start + yy * cl.ncx + x0);
      for (int xx = x0; xx <= x1; ++xx) {
        const int e = __ldg(cl.cell_start + yy * cl.ncx + xx + 1);
        const int a = lower_bound_i32(cl.zrank, s, e, qr.wlo);
        const int b = lower_bound_i32(cl.zrank, a, e, qr.whi);
        if (b > a) {
          run_lo[nrun][threadIdx.x] = a;
          run_hi[nrun][threadIdx.x] = b;
          ++nrun;
        }
        s = e;
      }
    }
    int c = 0;
    int j = nrun ? run_lo[0][threadIdx.x] : 0, end = nrun ? run_hi[0][threadIdx.x] : 0;
    while (c < nrun) {
      const double pra = __ldg(cl.ra + j), pdec = __ldg(cl.dec + j);
      if (pra >= qr.ra_lo && pra <= qr.ra_hi && pdec >= qr.dec_lo && pdec <= qr.dec_hi &&
          bt.inside(pra, pdec, __ldg(cl.cosdec + j)))
        ++count;
      if (++j == end && ++c < nrun) {
        j = run_lo[c][threadIdx.x];
        end = run_hi[c][threadIdx.x];
      }
    }
  }
  N_row[cl.row[q]] = count;
}

// ---- K6: candidate keys: N descending, ties by descending row (oracle pin D1 of fof.py:79-81) ----
__global__ void k_candidate_keys(int n, const int* N_row, int min_count, unsigned long long* key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int N = N_row[i];
  key[i] = N >= min_count ? (((unsigned long long)(0x7fffffff - N)) << 32) | (unsigned)(0x7fffffff - i) : ~0ull;
}

struct IsCandidate {
  __host__ __device__ bool operator()(const unsigned long long& k) const { return k != ~0ull; }
};

__global__ void k_keys_to_rows(int m, const unsigned long long* key, int* rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) rows[i] = 0x7fffffff - (int)(unsigned)(key[i] & 0xffffffffull);
}

void Stage0::run(fofb_handle* h, const fofb_params& p, Catalog& cat, CellList& cells) {
  cudaStream_t st = h->stream;
  const int n = (int)cat.n;
  const int B = 256;
  const Cosmo cosmo = device_cosmo(h, p);
  QueryRec* recs = rec.alloc(n);
  double* th = theta.alloc(n);
  CatalogView cv = cat.view();
  FOFB_PROFILED(h, "k_count_prep", k_count_prep<<<grid_for(n, B), B, 0, st>>>(cv, cosmo, p.count_radius, p.max_velocity, p.min_window, p.grid_cells,
                                             recs, th));
  // largest query radius sizes the cell list
  double* tmax = theta_max.alloc(1);
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceReduce::Max(nullptr, bytes, th, tmax, n, st));
  FOFB_CUDA(cub::DeviceReduce::Max(h->cub_tmp.alloc(bytes + 256), bytes, th, tmax, n, st));
  count_launch(h, 2);
  double r_max = 0.0;
  FOFB_CUDA(cudaMemcpyAsync(&r_max, tmax, sizeof(double), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  FOFB_REQUIRE(std::isfinite(r_max) && r_max > 0.0, "non-finite angular radius (check redshifts and cosmology)");
  cells.build(h, cat, r_max);
  int* N = N_row.alloc(n);
  FOFB_PROFILED(h, "k_count", k_count<<<grid_for(n, 256), 256, 0, st>>>(cells.view(), n, recs, N));

  unsigned long long* k_all = key_all.alloc(n);
  unsigned long long* k_sel = key_sel.alloc(n);
  unsigned long long* k_sorted = key_sorted.alloc(n);
  int* d_m = num_sel.alloc(1);
  k_candidate_keys<<<grid_for(n, B), B, 0, st>>>(n, N, p.min_count, k_all);
  FOFB_LAUNCH_CHECK(h);
  FOFB_CUDA(cub::DeviceSelect::If(nullptr, bytes, k_all, k_sel, d_m, n, IsCandidate(), st));
  FOFB_CUDA(cub::DeviceSelect::If(h->cub_tmp.alloc(bytes + 256), bytes, k_all, k_sel, d_m, n, IsCandidate(), st));
  count_launch(h, 2);
  int m = 0;
  FOFB_CUDA(cudaMemcpyAsync(&m, d_m, sizeof(int), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  n_candidates = m;
  int* rows = cand_rows.alloc(std::max(m, 1));
  if (m > 0) {
    FOFB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, k_sel, k_sorted, m, 0, 64, st));
    FOFB_CUDA(cub::DeviceRadixSort::SortKeys(h->cub_tmp.alloc(bytes + 256), bytes, k_sel, k_sorted, m, 0, 64, st));
    count_launch(h, 8);
    k_keys_to_rows<<<grid_for(m, B), B, 0, st>>>(m, k_sorted, rows);
    FOFB_LAUNCH_CHECK(h);
  }
  n_rows = n;
  valid = true;
}

}  // namespace fofb
