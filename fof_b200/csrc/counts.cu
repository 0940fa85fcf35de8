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
    const Box b = grid_box(bb, grid_cells, cat.zra[r], cat.zdec[r], theta);
    if (b.ra_lo != -INFINITY || b.ra_hi != INFINITY || b.dec_lo != -INFINITY || b.dec_hi != INFINITY) {
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
// Load balance.  A cluster member has a few hundred in-window candidates, a field galaxy a handful, and they sit in the
// same warps: a loop per galaxy leaves half the lanes idle.  So a batch of up to 256 galaxies first only LOCATES its
// runs (phase A: per galaxy the in-window run of each neighbour cell), and the (galaxy, candidate) pairs of the whole
// batch are then dealt out evenly to the threads of the CTA (phase B): thread t takes pairs [t C, (t + 1) C) of the
// batch's concatenated runs, finds its first galaxy by a binary search over the per-galaxy totals and walks on from
// there, adding its partial counts to the galaxies' shared-memory counters.
constexpr int kCtRunsPerQuery = 9;

struct CountBatch {
  float fx[kCtThreads], fy[kCtThreads], fcos[kCtThreads], T_lo[kCtThreads], T_hi[kCtThreads];
  int item_start[kCtThreads + 1];
  int count[kCtThreads];
  unsigned short run_first[kCtThreads * kCtRunsPerQuery], run_len[kCtThreads * kCtRunsPerQuery];
  unsigned char run_cell[kCtThreads * kCtRunsPerQuery];  // 3 * row + (cell - own cell + 1)
  unsigned char nrun[kCtThreads];
};

__global__ void __launch_bounds__(kCtThreads, 3) k_count(CellListView cl, int nsx, const CountRec* __restrict__ rec,
                                                        const double4* __restrict__ box, int* __restrict__ N_row) {
  extern __shared__ __align__(128) unsigned char ct_smem[];
  __shared__ int s_cs[3][kCtW + 3];
  __shared__ int s_delta[3];
  __shared__ int s_fits;
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ CountBatch B;
  typedef cub::BlockScan<int, kCtThreads> Scan;
  __shared__ typename Scan::TempStorage scan_tmp;
  Tile t;
  t.cs = s_cs;
  t.delta = s_delta;
  const int tid = threadIdx.x;
  const int y = blockIdx.x / nsx, x0 = (blockIdx.x % nsx) * kCtW;
  // the strip's own galaxies are the queries; an empty strip stages nothing
  const int xe = x0 + kCtW < cl.ncx ? x0 + kCtW : cl.ncx;
  const int q0 = __ldg(cl.cell_start + (size_t)y * cl.ncx + x0), q1 = __ldg(cl.cell_start + (size_t)y * cl.ncx + xe);
  if (q1 <= q0) return;
  tile_stage(cl, y, x0, t, reinterpret_cast<PackedPoint*>(ct_smem),
             reinterpret_cast<unsigned short*>(ct_smem + (size_t)kCtMaxPts * sizeof(PackedPoint)), &s_bar, &s_fits);
  const bool staged = s_fits != 0;
  const double cell_deg = fmax(cl.s_ra, cl.s_dec);
  const float sx = (float)cl.s_ra, sy = (float)cl.s_dec;
  const int r_lo = y > 0 ? 0 : 1, r_hi = y < cl.ncy - 1 ? 2 : 1;
  for (int qbase = q0; qbase < q1; qbase += kCtThreads) {
    // ---- phase A: this thread's galaxy -- its query, and where its candidates are ----
    const int qi = qbase + tid;
    int total = 0, nrun = 0, direct = 0;
    if (qi < q1) {
      const PackedPoint self = t.pk[qi + t.delta[1]];
      const CountRec qr = rec[self.zr];
      const int wlo = qr.wlo, whi = qr.whi & 0x7fffffff;
      if (whi > wlo) {
        const double ra = cl.ra[qi];
        const int qcx = cl.cx_of(ra);
        const int cx0 = qcx > 0 ? qcx - 1 : 0, cx1 = qcx < cl.ncx - 1 ? qcx + 1 : cl.ncx - 1;
        const bool boxed = qr.whi < 0;
        float T_lo, T_hi;
        tile_thresholds(qr.r_deg, cell_deg, &T_lo, &T_hi);
        if (boxed || !staged) {
          // the GriSPy cell box binds (rare), or the neighbourhood is too large for shared memory: the galaxy counts alone
          BubbleTest bt;
          bt.init(ra, cl.dec[qi], cl.cosdec[qi], qr.r_deg);
          Box bx{-INFINITY, INFINITY, -INFINITY, INFINITY};
          if (boxed) {
            const double4 b = box[self.zr];
            bx = Box{b.x, b.y, b.z, b.w};
          }
          TileQuery q;
          q.fx = self.fx; q.fy = self.fy; q.fcos = self.fcos; q.T_lo = T_lo; q.T_hi = T_hi;
          q.cx = qcx; q.cy = y; q.wlo = wlo; q.whi = whi;
          direct = tile_count(cl, t, y, q, r_lo, r_hi, cx0, cx1, bt, boxed, bx);
        } else {
          B.fx[tid] = self.fx; B.fy[tid] = self.fy; B.fcos[tid] = self.fcos; B.T_lo[tid] = T_lo; B.T_hi[tid] = T_hi;
          for (int r = r_lo; r <= r_hi; ++r) {
            const int d = t.delta[r];
            int s = t.cs[r][cx0 - t.xa] + d;
            for (int cx = cx0; cx <= cx1; ++cx) {
              const int e = t.cs[r][cx - t.xa + 1] + d;
              int first, last;
              tile_run(cl, t, y, r, cx, s, e, wlo, whi, &first, &last);
              if (last > first) {
                B.run_first[tid * kCtRunsPerQuery + nrun] = (unsigned short)first;
                B.run_len[tid * kCtRunsPerQuery + nrun] = (unsigned short)(last - first);
                B.run_cell[tid * kCtRunsPerQuery + nrun] = (unsigned char)(3 * r + (cx - qcx + 1));
                ++nrun;
                total += last - first;
              }
              s = e;
            }
          }
        }
      }
    }
    B.nrun[tid] = (unsigned char)nrun;
    B.count[tid] = direct;
    int start, N;
    Scan(scan_tmp).ExclusiveSum(total, start, N);
    B.item_start[tid] = start;
    if (tid == 0) B.item_start[kCtThreads] = N;
    __syncthreads();
    // ---- phase B: the batch's (galaxy, candidate) pairs, an equal share per thread ----
    const bool light = N < 24 * kCtThreads;  // a field-only batch: a handful of pairs per galaxy, nothing to balance
    if (light) {
      int cnt = 0;
      for (int rr = 0; rr < nrun; ++rr) {
        const int slot = tid * kCtRunsPerQuery + rr;
        const int code = B.run_cell[slot];
        const int r = code / 3, dx = code % 3 - 1;
        const float ox = (float)dx * sx - B.fx[tid], oy = (float)(r - 1) * sy - B.fy[tid];
        const float fc = B.fcos[tid], T_lo = B.T_lo[tid], T_hi = B.T_hi[tid];
        const int jb = B.run_first[slot], je = jb + B.run_len[slot];
        for (int j = jb; j < je; ++j) {
          const PackedPoint p = t.pk[j];
          const float hx = kHalfDeg2RadF * (p.fx + ox), hy = kHalfDeg2RadF * (p.fy + oy);
          const float sl = hx * (1.0f - hx * hx * (1.0f / 6.0f)), sd = hy * (1.0f - hy * hy * (1.0f / 6.0f));
          const float a = sd * sd + (fc * p.fcos) * (sl * sl);
          bool hit;
          if (a < T_lo) hit = true;
          else if (a > T_hi) hit = false;
          else {
            const int gj = j - t.delta[r];
            BubbleTest bt;
            bt.init(cl.ra[qi], cl.dec[qi], cl.cosdec[qi], rec[t.pk[qi + t.delta[1]].zr].r_deg);
            hit = bt.inside(cl.ra[gj], cl.dec[gj], cl.cosdec[gj]);
          }
          cnt += hit ? 1 : 0;
        }
      }
      if (cnt) B.count[tid] += cnt;
    }
    const int C = (N + kCtThreads - 1) / kCtThreads;
    const int i0 = tid * C, i1 = i0 + C < N ? i0 + C : N;
    if (!light && i0 < i1) {
      int lo = 0, hi = kCtThreads;  // last galaxy whose first pair is <= i0 (galaxies without pairs share a start: the last wins)
      while (hi - lo > 1) {
        const int m = (lo + hi) >> 1;
        if (B.item_start[m] <= i0) lo = m; else hi = m;
      }
      int g = lo;
      int rr = 0, pos = i0 - B.item_start[g];
      while (pos >= (int)B.run_len[g * kCtRunsPerQuery + rr]) {
        pos -= B.run_len[g * kCtRunsPerQuery + rr];
        ++rr;
      }
      int cnt = 0;
      int i = i0;
      while (i < i1) {
        // the run (g, rr) from pos on, as far as this thread's share goes
        const int slot = g * kCtRunsPerQuery + rr;
        const int len = B.run_len[slot];
        const int take = len - pos < i1 - i ? len - pos : i1 - i;
        const int code = B.run_cell[slot];
        const int r = code / 3, dx = code % 3 - 1;
        const float ox = (float)dx * sx - B.fx[g], oy = (float)(r - 1) * sy - B.fy[g];
        const float fc = B.fcos[g], T_lo = B.T_lo[g], T_hi = B.T_hi[g];
        const int jb = B.run_first[slot] + pos;
        for (int j = jb; j < jb + take; ++j) {
          const PackedPoint p = t.pk[j];
          const float hx = kHalfDeg2RadF * (p.fx + ox), hy = kHalfDeg2RadF * (p.fy + oy);
          const float sl = hx * (1.0f - hx * hx * (1.0f / 6.0f)), sd = hy * (1.0f - hy * hy * (1.0f / 6.0f));
          const float a = sd * sd + (fc * p.fcos) * (sl * sl);
          bool hit;
          if (a < T_lo) hit = true;
          else if (a > T_hi) hit = false;
          else {  // inside the band: fp64, from the full-precision arrays
            const int gq = qbase + g, gj = j - t.delta[r];
            BubbleTest bt;
            bt.init(cl.ra[gq], cl.dec[gq], cl.cosdec[gq], rec[t.pk[gq + t.delta[1]].zr].r_deg);
            hit = bt.inside(cl.ra[gj], cl.dec[gj], cl.cosdec[gj]);
          }
          cnt += hit ? 1 : 0;
        }
        i += take;
        pos += take;
        if (pos == len) {
          pos = 0;
          if (++rr == (int)B.nrun[g]) {  // on to the next galaxy that has pairs
            if (cnt) atomicAdd(&B.count[g], cnt);
            cnt = 0;
            rr = 0;
            do { ++g; } while (g < kCtThreads && B.nrun[g] == 0);
          }
        }
      }
      if (cnt && g < kCtThreads) atomicAdd(&B.count[g], cnt);
    }
    __syncthreads();
    if (qi < q1) N_row[cl.row[qi]] = B.count[tid];
    __syncthreads();  // the batch arrays are reused
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
