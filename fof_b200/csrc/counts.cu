// Stage 0: N(0.5) number counts for every galaxy and the candidate-centre ordering
// (kernels K5, K6 of DESIGN.md).  Replaces candidate_center_search's O(n^2) loop
// (fof.py:65-81) and find_number_count (helper.py:149-188).
#include <cub/cub.cuh>

#include "catalog_host.cuh"
#include "stage0.cuh"
#include "tile.cuh"

namespace fofb {

// ---- K5a: per-galaxy query record, one thread per z-rank ----------------------------------------
// theta_0.5(z_i), the velocity window of galaxy i, the bounding box of that window (= the bbox of
// the GriSPy grid find_number_count builds, helper.py:181) and the resulting cell-box bounds.  The box binds only
// when a grid edge falls inside the sliver (r, r / cos dec] next to the query: the record keeps a flag
// (bit 31 of whi) and the four bounds go to a side array that is touched for those queries alone.
__global__ void k_count_prep(CatalogView cat, Cosmo cosmo, double count_radius, double vmax, int min_window,
                             int grid_cells, CountRec* rec, double4* box, double* theta_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= cat.n) return;
  const double z = cat.zs[r];
  const double theta = cosmo_linear_to_angular_dist(cosmo, count_radius, z);
  int lo, hi;
  velocity_window(cat.zs, (int)cat.n, z, vmax, &lo, &hi);
  CountRec q;
  q.r_deg = theta;
  q.wlo = lo;
  q.whi = hi;
  if (hi - lo >= min_window) {
    const BBox bb = range_bbox(cat, lo, hi);
    // almost never binds: the shortcut decides that on two divisions; the exact routine (a dozen) runs for the rest
    bool bound;
    const Box b = grid_box_fast(bb, grid_cells, (double)grid_cells / ((bb.ra_max + 1.0e-6) - (bb.ra_min - 1.0e-6)),
                                (double)grid_cells / ((bb.dec_max + 1.0e-6) - (bb.dec_min - 1.0e-6)), cat.zra[r], cat.zdec[r],
                                theta, &bound);
    if (bound) {
      box[r] = make_double4(b.ra_lo, b.ra_hi, b.dec_lo, b.dec_hi);
      q.whi |= (int)0x80000000;
    }
  } else {
    q.whi = lo;  // empty window => N stays 0 (fof.py:71)
  }
  rec[r] = q;
  theta_out[r] = theta;
}

// ---- K5: the count -------------------------------------------------------------------------------------------------
// One CTA per strip of kCtW consecutive cells of one cell row.  The strip's neighbourhood -- the same cells plus one on
// either side, in the rows above, at and below -- is three contiguous slices of the cell-ordered SoA (ra, dec, cos dec,
// z-rank): they are brought into shared memory with bulk asynchronous copies (cp.async.bulk, one mbarrier), and every
// galaxy of the strip then runs its query against shared memory: per neighbour cell a binary search of its z-rank
// window (cells are z-rank sorted), then the cell-box and fp64 haversine tests on the in-window candidates.
// A neighbourhood that does not fit (percolating blobs) is read from global memory by the same code.
// (A variant that first located every galaxy's runs and then dealt the (galaxy, candidate) pairs of a 256-galaxy batch
// evenly over the CTA was measured at 6.5 ms against 3.8 ms for this one at 10 M galaxies: the extra shared-memory
// state cost a resident CTA per SM and two barriers per batch, more than the idle lanes it recovered.)
__global__ void __launch_bounds__(kCtThreads, kCtResident) k_count(CellListView cl, int nsx, const CountRec* __restrict__ rec,
                                                        const double4* __restrict__ box, int* __restrict__ N_row) {
  extern __shared__ __align__(128) unsigned char ct_smem[];
  __shared__ int s_cs[3][kCtW + 3];
  __shared__ int s_delta[3];
  __shared__ int s_fits;
  __shared__ __align__(8) unsigned long long s_bar;
  Tile t;
  t.cs = s_cs;
  t.delta = s_delta;
  const int y = blockIdx.x / nsx, x0 = (blockIdx.x % nsx) * kCtW;
  // the strip's own galaxies are the queries; an empty strip stages nothing
  const int xe = x0 + kCtW < cl.ncx ? x0 + kCtW : cl.ncx;
  const int q0 = __ldg(cl.cell_start + (size_t)y * cl.ncx + x0), q1 = __ldg(cl.cell_start + (size_t)y * cl.ncx + xe);
  if (q1 <= q0) return;
  tile_stage(cl, y, x0, t, reinterpret_cast<PackedPoint*>(ct_smem),
             reinterpret_cast<unsigned short*>(ct_smem + (size_t)kCtMaxPts * sizeof(PackedPoint)), &s_bar, &s_fits);
  const double cell_deg = fmax(cl.s_ra, cl.s_dec);
  for (int qi = q0 + (int)threadIdx.x; qi < q1; qi += kCtThreads) {
    const PackedPoint self = t.pk[qi + t.delta[1]];
    const CountRec qr = rec[self.zr];
    TileQuery q;
    q.wlo = qr.wlo;
    q.whi = qr.whi & 0x7fffffff;
    int count = 0;
    if (q.whi > q.wlo) {
      const double ra = cl.ra[qi], dec = cl.dec[qi];
      BubbleTest bt;
      bt.init(ra, dec, cl.cosdec[qi], qr.r_deg);
      Box bx{-INFINITY, INFINITY, -INFINITY, INFINITY};
      const bool boxed = qr.whi < 0;
      if (boxed) {
        const double4 b = box[self.zr];
        bx = Box{b.x, b.y, b.z, b.w};
      }
      q.fx = self.fx; q.fy = self.fy; q.fcos = self.fcos;
      q.cx = cl.cx_of(ra);
      q.cy = y;
      tile_thresholds(qr.r_deg, cell_deg, &q.T_lo, &q.T_hi);
      const int cx0 = q.cx > 0 ? q.cx - 1 : 0, cx1 = q.cx < cl.ncx - 1 ? q.cx + 1 : cl.ncx - 1;
      count = tile_count(cl, t, y, q, y > 0 ? 0 : 1, y < cl.ncy - 1 ? 2 : 1, cx0, cx1,
                         [&](double pra, double pdec, double pcos) { return bt.inside(pra, pdec, pcos); }, boxed, bx);
    }
    N_row[cl.row[qi]] = count;
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

void Stage0::run(fofb_handle* h, const fofb_params& p, Catalog& cat, CellList& cells) {
  cudaStream_t st = h->stream;
  const int n = (int)cat.n;
  const int B = 256;
  const Cosmo cosmo = device_cosmo(h, p);
  CountRec* recs = rec.alloc(n);
  double4* boxes = box.alloc(n);
  double* th = theta.alloc(n);
  CatalogView cv = cat.view();
  FOFB_PROFILED(h, "k_count_prep", k_count_prep<<<grid_for(n, B), B, 0, st>>>(cv, cosmo, p.count_radius, p.max_velocity, p.min_window, p.grid_cells,
                                             recs, boxes, th));
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
  cells.build(h, cat, r_max, true);
  int* N = N_row.alloc(n);
  {
    static bool attr_set[64] = {};
    if (!attr_set[h->device & 63]) {
      FOFB_CUDA(cudaFuncSetAttribute(k_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCtSmemBytes));
      attr_set[h->device & 63] = true;
    }
    const CellListView clv = cells.view();
    const int nsx = (clv.ncx + kCtW - 1) / kCtW;
    FOFB_PROFILED(h, "k_count", k_count<<<nsx * clv.ncy, kCtThreads, kCtSmemBytes, st>>>(clv, nsx, recs, boxes, N));
  }

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
