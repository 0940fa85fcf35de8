// Generic GriSPy replacement for the helper-level API: counts of `bubble_neighbors` queries on an arbitrary point set
// (helper.py:149-188 find_number_count, helper.py:207-216 center_overdensity).  Brute force, one warp per query: the
// helper functions are called with small point sets; the fast paths are fofb_number_counts / fofb_fof_*.
#include <cub/cub.cuh>

#include "bubble.cuh"

namespace fofb {

__global__ void k_bubble_mask_bbox(View P, int n, int use_window, double zc, double vmax, unsigned char* mask, double4* partial) {
  typedef cub::BlockReduce<double, 256> Reduce;
  __shared__ typename Reduce::TempStorage tmp;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool in = false;
  double ra = 0, dec = 0;
  if (i < n) {
    in = !use_window || fabs(redshift_to_velocity(P.at(i, 2), zc)) <= vmax;
    mask[i] = in ? 1 : 0;
    ra = P.at(i, 0);
    dec = P.at(i, 1);
  }
  const double a = Reduce(tmp).Reduce(in ? ra : INFINITY, cub::Min());
  __syncthreads();
  const double b = Reduce(tmp).Reduce(in ? ra : -INFINITY, cub::Max());
  __syncthreads();
  const double c = Reduce(tmp).Reduce(in ? dec : INFINITY, cub::Min());
  __syncthreads();
  const double d = Reduce(tmp).Reduce(in ? dec : -INFINITY, cub::Max());
  if (threadIdx.x == 0) partial[blockIdx.x] = make_double4(a, b, c, d);
}

__global__ void k_bubble_bbox_final(int nb, const double4* partial, double4* out) {
  double4 v = make_double4(INFINITY, -INFINITY, INFINITY, -INFINITY);
  for (int i = threadIdx.x; i < nb; i += 32) v = bbox_merge(v, partial[i]);
  for (int o = 16; o; o >>= 1) {
    double4 w;
    w.x = __shfl_xor_sync(0xffffffffu, v.x, o); w.y = __shfl_xor_sync(0xffffffffu, v.y, o);
    w.z = __shfl_xor_sync(0xffffffffu, v.z, o); w.w = __shfl_xor_sync(0xffffffffu, v.w, o);
    v = bbox_merge(v, w);
  }
  if (threadIdx.x == 0) *out = v;
}

__global__ void k_bubble_count(View P, int n, const unsigned char* mask, const double4* bbox, int q, const double* centres,
                               const double* radius, int grid_cells, int* counts) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= q) return;
  const double4 b4 = *bbox;
  const BBox bb{b4.x, b4.y, b4.z, b4.w};
  int cnt = 0;
  if (b4.x <= b4.y) {  // non-empty point set
    // centres farther than r outside the grid on any axis get nothing only if ALL centres are (App. B.1)
    bool all_oof = true;
    for (int t = lane; t < q; t += 32) {
      const double cr = centres[2 * t], cd = centres[2 * t + 1], r = radius[t];
      const bool oof = (cr - r > bb.ra_max + 1e-6) || (cr + r < bb.ra_min - 1e-6) || (cd - r > bb.dec_max + 1e-6) ||
                       (cd + r < bb.dec_min - 1e-6);
      if (!oof) all_oof = false;
    }
    all_oof = __all_sync(0xffffffffu, all_oof);
    if (!all_oof) {
      const double cra = centres[2 * w], cdec = centres[2 * w + 1], r = radius[w];
      const Box box = grid_box(bb, grid_cells, cra, cdec, r);
      BubbleTest bt;
      bt.init(cra, cdec, cos(__dmul_rn(cdec, kDeg2Rad)), r);
      for (int i = lane; i < n; i += 32) {
        if (!mask[i]) continue;
        const double pra = P.at(i, 0), pdec = P.at(i, 1);
        if (box.contains(pra, pdec) && bt.inside(pra, pdec, cos(__dmul_rn(pdec, kDeg2Rad)))) ++cnt;
      }
    }
  }
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) counts[w] = cnt;
}

void Bubble::count(fofb_handle* h, const fofb_params& p, const View& P, int q, const double* centres_host, const double* radius_host,
                   int use_window, double zc, int* counts_host) {
  cudaStream_t st = h->stream;
  const int n = (int)P.rows;
  FOFB_REQUIRE(P.cols >= (use_window ? 3 : 2), "bubble count: points need ra, dec (and z when a velocity window is requested)");
  if (q == 0) return;
  if (n == 0) {
    for (int i = 0; i < q; ++i) counts_host[i] = 0;
    return;
  }
  const int nb = grid_for(n, 256);
  unsigned char* m = mask.alloc(n);
  double4* part = partial.alloc(nb + 1);
  k_bubble_mask_bbox<<<nb, 256, 0, st>>>(P, n, use_window, zc, p.max_velocity, m, part);
  FOFB_LAUNCH_CHECK(h);
  k_bubble_bbox_final<<<1, 32, 0, st>>>(nb, part, part + nb);
  FOFB_LAUNCH_CHECK(h);
  double* c = cen.alloc((size_t)q * 3);
  FOFB_CUDA(cudaMemcpyAsync(c, centres_host, sizeof(double) * 2 * q, cudaMemcpyHostToDevice, st));
  FOFB_CUDA(cudaMemcpyAsync(c + 2 * (size_t)q, radius_host, sizeof(double) * q, cudaMemcpyHostToDevice, st));
  int* out = counts.alloc(q);
  FOFB_PROFILED(h, "k_bubble_count", k_bubble_count<<<grid_for((int64_t)q * 32, 128), 128, 0, st>>>(P, n, m, part + nb, q, c, c + 2 * (size_t)q, p.grid_cells, out));
  FOFB_CUDA(cudaMemcpyAsync(counts_host, out, sizeof(int) * q, cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
}

}  // namespace fofb
