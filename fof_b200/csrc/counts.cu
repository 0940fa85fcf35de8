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

// ---- K5: the count.  One warp per home cell; the three cell rows around it are staged through
// shared memory in chunks; every lane owns one home galaxy and scans the staged candidates
// (shared-memory broadcast reads), testing z-rank window -> cell box -> fp64 haversine.
constexpr int kCountWarps = 8;
constexpr int kCountChunk = 128;

__global__ void __launch_bounds__(32 * kCountWarps) k_count(CellListView cl, const QueryRec* __restrict__ rec,
                                                            int* __restrict__ N_row) {
  __shared__ int s_zr[kCountWarps][kCountChunk];
  __shared__ double s_ra[kCountWarps][kCountChunk];
  __shared__ double s_dec[kCountWarps][kCountChunk];
  __shared__ double s_cos[kCountWarps][kCountChunk];
  const int lane = threadIdx.x;
  const int w = threadIdx.y;
  const int ncell = cl.ncx * cl.ncy;
  for (int cell = blockIdx.x * kCountWarps + w; cell < ncell; cell += gridDim.x * kCountWarps) {
    const int h0 = cl.cell_start[cell], h1 = cl.cell_start[cell + 1];
    if (h0 == h1) continue;
    const int cy = cell / cl.ncx, cx = cell - cy * cl.ncx;
    const int x0 = cx > 0 ? cx - 1 : 0, x1 = cx < cl.ncx - 1 ? cx + 1 : cl.ncx - 1;
    const int y0 = cy > 0 ? cy - 1 : 0, y1 = cy < cl.ncy - 1 ? cy + 1 : cl.ncy - 1;
    for (int hb = h0; hb < h1; hb += 32) {
      const int q = hb + lane;
      const bool valid = q < h1;
      QueryRec qr;
      BubbleTest bt;
      unsigned wlo = 0, wlen = 0;
      if (valid) {
        qr = rec[cl.zrank[q]];
        bt.init(cl.ra[q], cl.dec[q], cl.cosdec[q], qr.r_deg);
        wlo = (unsigned)qr.wlo;
        wlen = (unsigned)(qr.whi - qr.wlo);
      }
      int count = 0;
      for (int yy = y0; yy <= y1; ++yy) {
        const int g0 = cl.cell_start[yy * cl.ncx + x0], g1 = cl.cell_start[yy * cl.ncx + x1 + 1];
        for (int c0 = g0; c0 < g1; c0 += kCountChunk) {
          const int m = min(kCountChunk, g1 - c0);
          __syncwarp();
          for (int j = lane; j < m; j += 32) {
            s_zr[w][j] = cl.zrank[c0 + j];
            s_ra[w][j] = cl.ra[c0 + j];
            s_dec[w][j] = cl.dec[c0 + j];
            s_cos[w][j] = cl.cosdec[c0 + j];
          }
          __syncwarp();
          if (wlen) {
            for (int j = 0; j < m; ++j) {
              if ((unsigned)s_zr[w][j] - wlo < wlen) {
                const double pra = s_ra[w][j], pdec = s_dec[w][j];
                if (pra >= qr.ra_lo && pra <= qr.ra_hi && pdec >= qr.dec_lo && pdec <= qr.dec_hi &&
                    bt.inside(pra, pdec, s_cos[w][j]))
                  ++count;
              }
            }
          }
        }
      }
      if (valid) N_row[cl.row[q]] = count;
    }
  }
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

void Stage0::run(fofb_handle* h, const fofb_params& p, Catalog& cat) {
  cudaStream_t st = h->stream;
  const int n = (int)cat.n;
  const int B = 256;
  const Cosmo cosmo = device_cosmo(h, p);
  QueryRec* recs = rec.alloc(n);
  double* th = theta.alloc(n);
  CatalogView cv = cat.view();
  k_count_prep<<<grid_for(n, B), B, 0, st>>>(cv, cosmo, p.count_radius, p.max_velocity, p.min_window, p.grid_cells,
                                             recs, th);
  FOFB_LAUNCH_CHECK(h);
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
  cat.build_cells(h, r_max);
  cv = cat.view();
  int* N = N_row.alloc(n);
  const int ncell = cat.ncx * cat.ncy;
  const int blocks = std::min((ncell + kCountWarps - 1) / kCountWarps, h->sm_count * 32);
  k_count<<<blocks, dim3(32, kCountWarps), 0, st>>>(cv.cells, recs, N);
  FOFB_LAUNCH_CHECK(h);

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
