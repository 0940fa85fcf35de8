// Nearest-cluster matching of a found-cluster sample against a published catalogue (catalog_matching.py:21-111):
// per catalogue row, the sample rows with |z_s - z_c| <= z_cut (:78-80), the nearest of them on the sky
// (match_coordinates_sky, :89-91), accepted when its Vincenty separation is within the angle of :85-88.
//
// Layout: the sample is sorted by z once, so the redshift cut of every catalogue row is one contiguous range found by
// binary search with the reference's own predicate; unit vectors are stored SoA in that order.  One warp per catalogue
// row scans its range (coalesced) and reduces (chord^2, row) lexicographically -- pinned order D5.
#include <cub/cub.cuh>

#include "geom.cuh"
#include "match.cuh"

namespace fofb {

__global__ void k_match_keys(View S, int ns, double* z, int* idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  z[i] = S.at(i, 2);
  idx[i] = i;
}

__global__ void k_match_vectors(View S, int ns, const int* __restrict__ row_sorted, double* x, double* y, double* z) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ns) return;
  const int r = row_sorted[j];
  const double lon = __dmul_rn(S.at(r, 0), kDeg2Rad), lat = __dmul_rn(S.at(r, 1), kDeg2Rad);
  double sl, cl, sb, cb;
  sincos(lon, &sl, &cl);
  sincos(lat, &sb, &cb);
  x[j] = __dmul_rn(cb, cl);
  y[j] = __dmul_rn(cb, sl);
  z[j] = sb;
}

// catalog_matching.py:85-88: proper length -> angle -> comoving length (x D_M) -> angle again, in degrees
__device__ double matching_angle_deg(const Cosmo& c, double d_mpc_h, double zc) {
  const double da = cosmo_angular_diameter_distance(c, zc);
  const double first_deg = __dmul_rn(__ddiv_rn(__ddiv_rn(d_mpc_h, c.h), da), kRad2Deg);
  const double comoving = __dmul_rn(__dmul_rn(first_deg, kDeg2Rad), cosmo_transverse_distance(c, zc));
  return __dmul_rn(__ddiv_rn(comoving, da), kRad2Deg);
}

__global__ void __launch_bounds__(128) k_match(View S, int ns, View Cat, int nc, const double* __restrict__ zs,
                                               const int* __restrict__ row_sorted, const double* __restrict__ x,
                                               const double* __restrict__ y, const double* __restrict__ z,
                                               const double* __restrict__ zrange, Cosmo cosmo, double d_mpc_h, double z_cut,
                                               int* match_row, double* match_sep, int* counter) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= nc) return;
  const double cra = Cat.at(w, 0), cdec = Cat.at(w, 1), cz = Cat.at(w, 2);
  // :66-68 keep catalogue rows with min_z - 0.1 <= z <= max_z + 0.1 of the sample
  if (!(cz >= __dsub_rn(zrange[0], 0.1) && cz <= __dadd_rn(zrange[1], 0.1))) {
    if (lane == 0) { match_row[w] = -2; match_sep[w] = nan(""); }
    return;
  }
  // |fl(z_s - z_c)| <= z_cut is monotone in z_s: first index with fl(z_s - z_c) >= -z_cut, first with > z_cut
  int lo = 0, hi = ns;
  while (lo < hi) { const int m = (lo + hi) >> 1; if (__dsub_rn(zs[m], cz) >= -z_cut) hi = m; else lo = m + 1; }
  const int first = lo;
  hi = ns;
  while (lo < hi) { const int m = (lo + hi) >> 1; if (__dsub_rn(zs[m], cz) > z_cut) hi = m; else lo = m + 1; }
  const int last = lo;
  const double lon = __dmul_rn(cra, kDeg2Rad), lat = __dmul_rn(cdec, kDeg2Rad);
  double sl, cl, sb, cb;
  sincos(lon, &sl, &cl);
  sincos(lat, &sb, &cb);
  const double cx = __dmul_rn(cb, cl), cy = __dmul_rn(cb, sl), czz = sb;
  double best = INFINITY;
  int best_row = 0x7fffffff;
  for (int j = first + lane; j < last; j += 32) {
    const double dx = __dsub_rn(x[j], cx), dy = __dsub_rn(y[j], cy), dz = __dsub_rn(z[j], czz);
    const double c2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    const int r = row_sorted[j];
    if (c2 < best || (c2 == best && r < best_row)) { best = c2; best_row = r; }
  }
  for (int o = 16; o; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int orow = __shfl_xor_sync(0xffffffffu, best_row, o);
    if (ob < best || (ob == best && orow < best_row)) { best = ob; best_row = orow; }
  }
  if (lane) return;
  if (first >= last) { match_row[w] = -1; match_sep[w] = nan(""); return; }
  // :90 d2d = catalogcoord[idx].separation(matchcoord): Vincenty from the sample row to the catalogue row
  const double sep = __dmul_rn(vincenty_rad(__dmul_rn(S.at(best_row, 0), kDeg2Rad), __dmul_rn(S.at(best_row, 1), kDeg2Rad), lon, lat),
                               kRad2Deg);
  match_sep[w] = sep;
  if (sep <= matching_angle_deg(cosmo, d_mpc_h, cz)) {
    match_row[w] = best_row;
    atomicAdd(counter, 1);
  } else {
    match_row[w] = -1;
  }
}

void Matcher::run(fofb_handle* h, const fofb_params& p, const View& S, const View& Cat, double d_mpc_h, double z_cut) {
  cudaStream_t st = h->stream;
  const int ns = (int)S.rows, nc = (int)Cat.rows;
  FOFB_REQUIRE(S.cols >= 3 && Cat.cols >= 3, "catalogue matching: sample and catalogue need ra, dec, z columns");
  FOFB_REQUIRE(ns > 0, "catalogue matching: the sample is empty (the reference raises on min() of an empty array)");
  n_matched = 0;
  if (nc == 0) return;
  double* zi = z_in.alloc(ns);
  int* ii = idx_in.alloc(ns);
  double* zsrt = z_sorted.alloc(ns);
  int* rs = row_sorted.alloc(ns);
  k_match_keys<<<grid_for(ns, 256), 256, 0, st>>>(S, ns, zi, ii);
  FOFB_LAUNCH_CHECK(h);
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, zi, zsrt, ii, rs, ns, 0, 64, st));
  FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.alloc(bytes), bytes, zi, zsrt, ii, rs, ns, 0, 64, st));
  h->launches += 3;
  double* zr = zrange.alloc(2);  // min = first, max = last of the sorted keys
  FOFB_CUDA(cudaMemcpyAsync(zr, zsrt, sizeof(double), cudaMemcpyDeviceToDevice, st));
  FOFB_CUDA(cudaMemcpyAsync(zr + 1, zsrt + ns - 1, sizeof(double), cudaMemcpyDeviceToDevice, st));
  double* vx = x.alloc(ns); double* vy = y.alloc(ns); double* vz = z.alloc(ns);
  k_match_vectors<<<grid_for(ns, 256), 256, 0, st>>>(S, ns, rs, vx, vy, vz);
  FOFB_LAUNCH_CHECK(h);
  int* mr = match_row.alloc(nc);
  double* ms = match_sep.alloc(nc);
  int* cnt = counter.alloc(1);
  FOFB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int), st));
  const Cosmo cosmo = device_cosmo(h, p);
  FOFB_PROFILED(h, "k_match", k_match<<<grid_for((int64_t)nc * 32, 128), 128, 0, st>>>(S, ns, Cat, nc, zsrt, rs, vx, vy, vz, zr, cosmo,
                                                                                        d_mpc_h, z_cut, mr, ms, cnt));
  int c = 0;
  FOFB_CUDA(cudaMemcpyAsync(&c, cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  n_matched = c;
}

}  // namespace fofb
