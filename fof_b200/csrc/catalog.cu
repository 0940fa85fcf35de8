// Catalogue build: z sort, range-bbox tables, (dec, ra) cell list (kernels K1-K4 of DESIGN.md).
#include <cub/cub.cuh>

#include "catalog_host.cuh"

namespace fofb {

// ---- K1a: pull (ra, dec, z) out of the strided row matrix; NaN rows are refused later ----------
__global__ void k_extract(View rows, int n, double* ra, double* dec, double* z, int* idx, int* unsorted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i > 0) {  // the reference's pipeline hands over rows ordered by redshift (pipeline.py:47,56): detect that
    const double prev = rows.at(i - 1, 2);
    if (!((isfinite(prev) ? prev : INFINITY) <= (isfinite(rows.at(i, 2)) ? rows.at(i, 2) : INFINITY))) *unsorted = 1;
  }
  // NaN would slip through the min/max reductions of the bounding box: store +inf so that Catalog::build refuses it
  const double a = rows.at(i, 0), d = rows.at(i, 1), zz = rows.at(i, 2);
  ra[i] = isfinite(a) ? a : INFINITY;
  dec[i] = isfinite(d) ? d : INFINITY;
  z[i] = isfinite(zz) ? zz : INFINITY;
  idx[i] = i;
}

// ---- after the z sort: z-ordered ra/dec, rank of each row ------------------------------------------
__global__ void k_zorder_gather(int n, const int* zorder, const double* ra, const double* dec, double* zra,
                                double* zdec, int* rank) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int row = zorder[r];
  zra[r] = ra[row];
  zdec[r] = dec[row];
  rank[row] = r;
}

// ---- range-bbox level 0: per-block inclusive prefix / suffix boxes and the block aggregate ---------
__global__ void __launch_bounds__(kBBoxBlock) k_bbox_blocks(int n, const double* zra, const double* zdec,
                                                            double4* pre, double4* suf, double4* level0) {
  typedef cub::BlockScan<double4, kBBoxBlock> Scan;
  __shared__ typename Scan::TempStorage tmp;
  struct Merge {
    __device__ double4 operator()(const double4& a, const double4& b) const { return bbox_merge(a, b); }
  };
  const int base = blockIdx.x * kBBoxBlock;
  const int i = base + threadIdx.x;
  const double4 ident = make_double4(INFINITY, -INFINITY, INFINITY, -INFINITY);
  double4 v = ident;
  if (i < n) v = make_double4(zra[i], zra[i], zdec[i], zdec[i]);
  double4 p, agg;
  Scan(tmp).InclusiveScan(v, p, Merge(), agg);
  if (i < n) pre[i] = p;
  if (threadIdx.x == 0) level0[blockIdx.x] = agg;
  __syncthreads();
  // suffix = scan of the reversed block
  const int j = base + (kBBoxBlock - 1 - threadIdx.x);
  double4 w = ident;
  if (j < n) w = make_double4(zra[j], zra[j], zdec[j], zdec[j]);
  double4 s;
  Scan(tmp).InclusiveScan(w, s, Merge());
  if (j < n) suf[j] = s;
}

__global__ void k_bbox_level(int nb, int k, const double4* prev, double4* next) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const int j = i + (1 << (k - 1));
  next[i] = j < nb ? bbox_merge(prev[i], prev[j]) : prev[i];
}

// ---- K1b: cell key of every galaxy, in z-rank order ---------------------------------------------------
__global__ void k_cell_keys(int n, const double* zra, const double* zdec, CellListView cl, unsigned* key, int* val) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  key[r] = (unsigned)(cl.cy_of(zdec[r]) * cl.ncx + cl.cx_of(zra[r]));
  val[r] = r;
}

// ---- K3: gather into the cell-ordered SoA ---------------------------------------------------------------
__global__ void k_cell_gather(int n, const int* sorted_rank, const int* zorder, const double* zra, const double* zdec,
                              double* ra, double* dec, double* cosdec, int* zrank, int* row, CellListView cl, PackedPoint* pk) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int r = sorted_rank[k];
  const double a = zra[r], d = zdec[r];
  const double c = cos(__dmul_rn(d, kDeg2Rad));
  ra[k] = a;
  dec[k] = d;
  cosdec[k] = c;
  zrank[k] = r;
  row[k] = zorder[r];
  if (pk) {
    PackedPoint p;
    p.fx = (float)(a - (cl.ra0 + (double)cl.cx_of(a) * cl.s_ra));
    p.fy = (float)(d - (cl.dec0 + (double)cl.cy_of(d) * cl.s_dec));
    p.fcos = (float)c;
    p.zr = r;
    pk[k] = p;
  }
}

// ---- K4: cell table from the sorted keys ----------------------------------------------------------------
__global__ void k_cell_starts(int n, const unsigned* key, int ncell, int* cell_start) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n) return;
  const int cur = k < n ? (int)key[k] : ncell;
  const int prev = k > 0 ? (int)key[k - 1] : -1;
  for (int c = prev + 1; c <= cur; ++c) cell_start[c] = k;
}

// ---- z-rank bins per cell: start offsets of every bin, from the cell-ordered z-ranks ------------------------------------
__global__ void k_ztab(int n, const unsigned* key, const int* cell_start, const int* zrank, int zshift, unsigned short* tab,
                       int* too_long) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int c = (int)key[k];
  const int s = cell_start[c], e = cell_start[c + 1];
  if (e - s >= 65535) {
    if (k == s) *too_long = 1;
    return;
  }
  const int b = zrank[k] >> zshift;
  const int pb = k == s ? -1 : (zrank[k - 1] >> zshift);
  unsigned short* t = tab + (size_t)c * kZBins;
  for (int bb = pb + 1; bb <= b; ++bb) t[bb] = (unsigned short)(k - s);
  if (k + 1 == e)
    for (int bb = b + 1; bb < kZBins; ++bb) t[bb] = (unsigned short)(e - s);
}

struct MinMaxOp {
  __host__ __device__ double4 operator()(const double4& a, const double4& b) const {
    return make_double4(a.x < b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z < b.z ? a.z : b.z, a.w > b.w ? a.w : b.w);
  }
};

static unsigned char* cub_tmp(fofb_handle* h, size_t bytes) { return h->cub_tmp.alloc(bytes + 256); }

void Catalog::build(fofb_handle* h, const View& rows_in) {
  rows = rows_in;
  n = rows.rows;
  FOFB_REQUIRE(n > 0 && n < (int64_t)2000000000, "catalogue must have between 1 and 2e9 rows");
  FOFB_REQUIRE(rows.cols >= 3, "catalogue needs at least the ra, dec, z columns");
  cudaStream_t st = h->stream;
  const int ni = (int)n;
  const int B = 256;
  double* ra = ra_row.alloc(n);
  double* dec = dec_row.alloc(n);
  double* z = z_row.alloc(n);
  int* idx = idx_tmp.alloc(n);
  int* unsorted = sorted_flag.alloc(1);
  FOFB_CUDA(cudaMemsetAsync(unsorted, 0, sizeof(int), st));
  FOFB_PROFILED(h, "k_extract", k_extract<<<grid_for(n, B), B, 0, st>>>(rows, ni, ra, dec, z, idx, unsorted));
  int host_unsorted = 0;
  FOFB_CUDA(cudaMemcpyAsync(&host_unsorted, unsorted, sizeof(int), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));

  // K2a: stable radix sort by redshift -- the identity when the rows already come in redshift order
  double* zs_p = zs.alloc(n);
  int* zorder_p = zorder.alloc(n);
  size_t bytes = 0;
  if (host_unsorted) {
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, z, zs_p, idx, zorder_p, ni, 0, 64, st));
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, z, zs_p, idx, zorder_p, ni, 0, 64, st));
    count_launch(h, 8);
  } else {
    FOFB_CUDA(cudaMemcpyAsync(zs_p, z, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    FOFB_CUDA(cudaMemcpyAsync(zorder_p, idx, sizeof(int) * n, cudaMemcpyDeviceToDevice, st));
  }
  double* zra_p = zra.alloc(n);
  double* zdec_p = zdec.alloc(n);
  int* rank_p = rank.alloc(n);
  k_zorder_gather<<<grid_for(n, B), B, 0, st>>>(ni, zorder_p, ra, dec, zra_p, zdec_p, rank_p);
  FOFB_LAUNCH_CHECK(h);

  // range-bbox tables
  nblocks = (ni + kBBoxBlock - 1) / kBBoxBlock;
  int nlev = 1;
  while ((1 << nlev) <= nblocks) ++nlev;
  double4* pre_p = pre.alloc(n);
  double4* suf_p = suf.alloc(n);
  double4* lev_p = levels.alloc((size_t)nlev * nblocks);
  FOFB_PROFILED(h, "k_bbox_blocks", k_bbox_blocks<<<nblocks, kBBoxBlock, 0, st>>>(ni, zra_p, zdec_p, pre_p, suf_p, lev_p));
  for (int k = 1; k < nlev; ++k) {
    k_bbox_level<<<grid_for(nblocks, B), B, 0, st>>>(nblocks, k, lev_p + (size_t)(k - 1) * nblocks, lev_p + (size_t)k * nblocks);
    FOFB_LAUNCH_CHECK(h);
  }
  // global bbox = reduction of level 0
  double4* gb = gbox_dev.alloc(1);
  const double4 ident = make_double4(INFINITY, -INFINITY, INFINITY, -INFINITY);
  FOFB_CUDA(cub::DeviceReduce::Reduce(nullptr, bytes, lev_p, gb, nblocks, MinMaxOp(), ident, st));
  FOFB_CUDA(cub::DeviceReduce::Reduce(cub_tmp(h, bytes), bytes, lev_p, gb, nblocks, MinMaxOp(), ident, st));
  count_launch(h, 2);
  double4 g;
  FOFB_CUDA(cudaMemcpyAsync(&g, gb, sizeof(g), cudaMemcpyDeviceToHost, st));
  double zmm[2];
  FOFB_CUDA(cudaMemcpyAsync(&zmm[0], zs_p, sizeof(double), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaMemcpyAsync(&zmm[1], zs_p + (n - 1), sizeof(double), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  gbox = BBox{g.x, g.y, g.z, g.w};
  zmin = zmm[0];
  zmax = zmm[1];
  FOFB_REQUIRE(std::isfinite(g.x) && std::isfinite(g.y) && std::isfinite(g.z) && std::isfinite(g.w) &&
                   std::isfinite(zmin) && std::isfinite(zmax),
               "catalogue contains non-finite ra/dec/z values");
  FOFB_REQUIRE(g.z >= -90.0 && g.w <= 90.0, "declination outside [-90, 90] degrees");
  dec_absmax = fmax(fabs(g.z), fabs(g.w));
}

void CellList::build(fofb_handle* h, const Catalog& cat, double r_max_deg, bool packed) {
  FOFB_REQUIRE(r_max_deg > 0.0 && std::isfinite(r_max_deg), "cell list needs a positive finite query radius");
  cudaStream_t st = h->stream;
  const int64_t n = cat.n;
  const int ni = (int)n;
  const int B = 256;
  s_dec = r_max_deg * 1.000001;
  s_ra = ra_reach(r_max_deg, cat.dec_absmax);
  // keep the table within ~2 cells per galaxy (and 2^30 in total): larger cells are always valid
  const double ext_ra = cat.gbox.ra_max - cat.gbox.ra_min, ext_dec = cat.gbox.dec_max - cat.gbox.dec_min;
  for (;;) {
    const double cells = (floor(ext_ra / s_ra) + 1.0) * (floor(ext_dec / s_dec) + 1.0);
    if (cells <= fmax(1024.0, 2.0 * (double)n) && cells < 1.0e9) break;
    s_ra *= 1.25;
    s_dec *= 1.25;
  }
  ncx = (int)floor(ext_ra / s_ra) + 1;
  ncy = (int)floor(ext_dec / s_dec) + 1;
  ra0 = cat.gbox.ra_min;
  dec0 = cat.gbox.dec_min;
  const int ncell = ncx * ncy;
  CellListView cl = view();

  unsigned* key_in = key_a.alloc(n);
  unsigned* key_out = key_b.alloc(n);
  int* val_in = val_a.alloc(n);
  int* val_out = val_b.alloc(n);
  k_cell_keys<<<grid_for(n, B), B, 0, st>>>(ni, cat.zra.p, cat.zdec.p, cl, key_in, val_in);
  FOFB_LAUNCH_CHECK(h);
  int bits = 1;
  while ((1ll << bits) < ncell) ++bits;
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in, key_out, val_in, val_out, ni, 0, bits, st));
  FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, key_in, key_out, val_in, val_out, ni, 0, bits, st));
  count_launch(h, 2 + (bits + 7) / 8);
  has_packed = packed;
  FOFB_PROFILED(h, "k_cell_gather", k_cell_gather<<<grid_for(n, B), B, 0, st>>>(ni, val_out, cat.zorder.p, cat.zra.p, cat.zdec.p, ra.alloc(n), dec.alloc(n),
                                              cosdec.alloc(n), zrank.alloc(n), row.alloc(n), cl, packed ? pk.alloc(n) : nullptr));
  int* cs = cell_start.alloc((size_t)ncell + 1);
  k_cell_starts<<<grid_for(n + 1, B), B, 0, st>>>(ni, key_out, ncell, cs);
  FOFB_LAUNCH_CHECK(h);
  has_ztab = false;
  if (packed) {
    zshift = 0;
    while (((ni - 1) >> zshift) >= kZBins) ++zshift;
    unsigned short* zt = ztab.alloc((size_t)ncell * kZBins + 64);
    int* flag = ztab_flag.alloc(1);
    FOFB_CUDA(cudaMemsetAsync(zt, 0, sizeof(unsigned short) * (size_t)ncell * kZBins, st));
    FOFB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
    k_ztab<<<grid_for(n, B), B, 0, st>>>(ni, key_out, cs, zrank.p, zshift, zt, flag);
    FOFB_LAUNCH_CHECK(h);
    int too_long = 0;
    FOFB_CUDA(cudaMemcpyAsync(&too_long, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    FOFB_CUDA(cudaStreamSynchronize(st));
    has_ztab = too_long == 0;
  }
  built = true;
  radius = r_max_deg;
}

CellListView CellList::view() const {
  CellListView v{};
  v.ra0 = ra0;
  v.dec0 = dec0;
  v.inv_sra = 1.0 / s_ra;
  v.inv_sdec = 1.0 / s_dec;
  v.s_ra = s_ra;
  v.s_dec = s_dec;
  v.pk = has_packed ? pk.p : nullptr;
  v.ztab = has_packed && has_ztab ? ztab.p : nullptr;
  v.zshift = zshift;
  v.ncx = ncx;
  v.ncy = ncy;
  v.cell_start = cell_start.p;
  v.ra = ra.p;
  v.dec = dec.p;
  v.cosdec = cosdec.p;
  v.zrank = zrank.p;
  v.row = row.p;
  return v;
}

CatalogView Catalog::view() const {
  CatalogView v{};
  v.n = n;
  v.zs = zs.p;
  v.zra = zra.p;
  v.zdec = zdec.p;
  v.zorder = zorder.p;
  v.rank = rank.p;
  v.pre = pre.p;
  v.suf = suf.p;
  v.levels = levels.p;
  v.nblocks = nblocks;
  v.dec_absmax = dec_absmax;
  return v;
}

}  // namespace fofb
