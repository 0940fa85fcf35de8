// Whole-job path: pipeline.py:62-64,80-99,113,173 (candidate search -> z cuts -> FoF -> interloper
// removal -> virial masses) with every intermediate kept on the device.  This is what bench.py times.
#include <cub/cub.cuh>

#include <ctime>

#include "catalog_host.cuh"
#include "clean.cuh"
#include "fof_stage.cuh"
#include "pipeline.cuh"
#include "stage0.cuh"

namespace fofb {

void fof_overlap_and_bcg(fofb_handle* h, FofState& S);  // overlap.cu

static unsigned char* cub_tmp(fofb_handle* h, size_t bytes) { return h->cub_tmp.alloc(bytes + 256); }


__global__ void k_scan_total_pl(const long long* len, long long* off, int n) { off[n] = off[n - 1] + len[n - 1]; }

static long long scan_offsets(fofb_handle* h, const long long* len, long long* off, int n) {
  if (n == 0) return 0;
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, len, off, n, h->stream));
  FOFB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp(h, bytes), bytes, len, off, n, h->stream));
  k_scan_total_pl<<<1, 1, 0, h->stream>>>(len, off, n);  // off[n] = total, on the device
  count_launch(h, 3);
  return d2h_scalar(h, off + n);
}

// pipeline.py:80-86: galaxies with zmin <= z <= zmax_gal keep their place in the FoF() input
__global__ void k_zflag(int n, const double* z, double zmin, double zmax, int* flag, int* iota) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = z[i];
  flag[i] = (v >= zmin && v <= zmax) ? 1 : 0;
  iota[i] = i;
}

__global__ void k_cand_flag(int m, const int* cand_rows, const double* z, double zmin, double zmax, double zlt, int* flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double v = z[cand_rows[i]];
  flag[i] = (v >= zmin && v <= zmax && v < zlt) ? 1 : 0;
}

// main_arr rows (fof.py:61): the input columns plus N(0.5), written row-major with 8 columns
__global__ void k_build_rows(int cnt, const int* src_rows, View raw, const int* N_row, double* out, int ocols) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)cnt * ocols) return;
  const int i = (int)(t / ocols), c = (int)(t % ocols);
  const int row = src_rows[i];
  out[t] = c < raw.cols ? raw.at(row, c) : (double)N_row[row];
}

__global__ void k_gather_members(long long total, const int* rows, const double* G, int cols, double* M) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total * cols) return;
  const long long i = t / cols;
  const int c = (int)(t % cols);
  M[t] = G[(long long)rows[i] * cols + c];
}

__global__ void k_cluster_centres(int K, const int* centre, const double* c_ra, const double* c_dec, const double* c_z,
                                  double* cen) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int c = centre[k];
  cen[3 * k] = c_ra[c];
  cen[3 * k + 1] = c_dec[c];
  cen[3 * k + 2] = c_z[c];
}

__global__ void k_rich_flag(int K, const long long* off, int richness, int* flag, int* iota) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  flag[k] = (off[k + 1] - off[k]) >= richness ? 1 : 0;
  iota[k] = k;
}

__global__ void k_sel_sizes(int K2, const int* sel, const long long* off, long long* len) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K2) len[k] = off[sel[k] + 1] - off[sel[k]];
}

// cleaned member rows of the kept clusters: position inside the cluster -> G row -> original row
__global__ void k_cleaned_rows(int K2, const int* sel, const long long* in_off, const int* in_rows,
                               const long long* kept_off, const int* kept_idx, const long long* out_off, int* out_grow) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= K2) return;
  const int s = sel[k];
  const long long src = kept_off[s], len = kept_off[s + 1] - src, base = in_off[s], dst = out_off[k];
  for (long long t = lane; t < len; t += 32) out_grow[dst + t] = in_rows[base + kept_idx[src + t]];
}

__global__ void k_final_meta(int K2, const int* sel, const int* centre, const int* N, const double* D, const int* cen_src,
                             const int* g_src, int m, int* out_centre_row, double* out_N, double* out_D) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K2) return;
  const int s = sel[k];
  const int c = centre[s];
  // distributed runs (cen_src == nullptr) report the index into the all-process centre table instead of a row
  out_centre_row[k] = !cen_src ? c : (c < m ? cen_src[c] : g_src[c - m]);
  out_N[k] = (double)N[s];
  out_D[k] = D[s];
}

__global__ void k_map_rows(long long total, const int* grow, const int* g_src, int* out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) out[t] = g_src[grow[t]];
}

__global__ void k_max_len(int K, const long long* off, int* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) atomicMax(out, (int)(off[k + 1] - off[k]));
}

void Pipeline::begin(fofb_handle* h, const fofb_params& p, const View& raw, double zmin, double zmax_gal, double zmax_cen,
                     const unsigned* mt_key, int mt_pos) {
  dist = false;
  fof.K_glob_merged = -1;
  fof.have_global_bbox = false;
  stage0_and_inputs(h, p, raw, zmin, zmax_gal, zmin, zmax_cen, mt_key, mt_pos);
  fof.run(h, p, Gv, Cv);
  fof.last_K = fof.merged.K;
  began = true;
}

// Parameter sweep: stage 0, the catalogue, the cell lists and the per-centre scalars are reused; only
// the priority rounds and what follows run again.
void Pipeline::rerun(fofb_handle* h, double linking_length_factor, int richness, double overdensity, const unsigned* mt_key,
                     int mt_pos) {
  if (!began || dist) throw Error(FOFB_ERR_STATE, "fofb_pipeline_rerun needs a completed single-process fofb_pipeline_begin / run");
  finished = false;
  params.linking_length_factor = linking_length_factor;
  params.richness = richness;
  params.overdensity = overdensity;
  if (mt_key) {
    fof.params_n_random_hint = params.n_random;
    fof.mt_prefetch(h, mt_key, mt_pos, (long long)(fof.last_K * 1.25) + 256);
  }
  fof.restart_rounds(h, linking_length_factor, richness, overdensity);
  while (fof.na > 0) {
    fof.round_evaluate(h);
    fof.round_advance(h);
  }
  fof.finalize(h);
  fof_overlap_and_bcg(h, fof);
  fof.ran = true;
  fof.last_K = fof.merged.K;
}

__global__ void k_cluster_sizes_ids(int K, const long long* off, const int* rows, const double* G, int cols, long long* sizes,
                                    double* ids) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  const long long s0 = off[k], len = off[k + 1] - s0;
  if (lane == 0) sizes[k] = len;
  for (long long t = lane; t < len; t += 32) ids[s0 + t] = G[(long long)rows[s0 + t] * cols + 6];
}

void Pipeline::export_local_clusters(fofb_handle* h, int* centre, long long* sizes, double* ids) {
  const ClusterList& L = fof.candidates;
  if (L.K == 0) return;
  cudaStream_t st = h->stream;
  FOFB_CUDA(cudaMemcpyAsync(centre, L.centre.p, sizeof(int) * L.K, cudaMemcpyDeviceToDevice, st));
  k_cluster_sizes_ids<<<grid_for((int64_t)L.K * 32, 128), 128, 0, st>>>(L.K, L.off.p, L.rows.p, g_rows.p, 8, sizes, ids);
  FOFB_LAUNCH_CHECK(h);
  FOFB_CUDA(cudaStreamSynchronize(st));
}

// Distributed variant: this process holds a redshift slab plus halos.  Rows with z in [valid_lo, valid_hi] have
// their whole velocity window on this process (so N(0.5) is exact for them and they may enter FoF()); candidate
// centres with own_lo <= z < own_hi are the ones this process evaluates.
void Pipeline::begin_dist(fofb_handle* h, const fofb_params& p, const View& raw, double zmin, double zmax_gal, double zmax_cen,
                          double valid_lo, double valid_hi, double own_lo, double own_hi) {
  dist = true;
  fof.K_glob_merged = -1;
  fof.have_global_bbox = false;
  own_hi_exclusive = own_hi;
  stage0_and_inputs(h, p, raw, fmax(zmin, valid_lo), fmin(zmax_gal, valid_hi), fmax(zmin, own_lo), zmax_cen, nullptr, 0);
  began = false;  // becomes true after the distributed rounds
}

void Pipeline::stage0_and_inputs(fofb_handle* h, const fofb_params& p, const View& raw, double zmin, double zmax_gal,
                                 double czmin, double zmax_cen, const unsigned* mt_key, int mt_pos) {
  cudaStream_t st = h->stream;
  const int B = 256;
  began = finished = false;
  params = p;
  if (mt_key) {  // overlap the serial MT19937 stream with stages 0-3
    fof.params_n_random_hint = p.n_random;
    const long long k_spec = fof.last_K > 0 ? (long long)(fof.last_K * 1.25) + 256 : raw.rows / 150 + 1024;
    fof.mt_prefetch(h, mt_key, mt_pos, k_spec);
  }
  FOFB_REQUIRE(raw.cols == 7, "the pipeline expects the 7 input columns ra, dec, z, z_lower, z_upper, RMag, ID");
  const bool trace = getenv("FOFB_TRACE") != nullptr;
  timespec t_a, t_b;
  if (trace) { cudaStreamSynchronize(st); clock_gettime(CLOCK_MONOTONIC, &t_a); }
  cat.build(h, raw);
  s0.run(h, p, cat, cells);
  if (trace) {
    cudaStreamSynchronize(st);
    clock_gettime(CLOCK_MONOTONIC, &t_b);
    fprintf(stderr, "[trace] stage 0 %.2f ms\n", (t_b.tv_sec - t_a.tv_sec) * 1e3 + (t_b.tv_nsec - t_a.tv_nsec) * 1e-6);
  }
  const int n = (int)cat.n;
  n_in = n;
  // galaxies that enter FoF()
  int* flag = flags.alloc(n);
  int* iota = iotas.alloc(n);
  int* dcount = counter.alloc(2);
  k_zflag<<<grid_for(n, B), B, 0, st>>>(n, cat.z_row.p, zmin, zmax_gal, flag, iota);
  FOFB_LAUNCH_CHECK(h);
  int* gsrc = g_src.alloc(n);
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, iota, flag, gsrc, dcount, n, st));
  FOFB_CUDA(cub::DeviceSelect::Flagged(cub_tmp(h, bytes), bytes, iota, flag, gsrc, dcount, n, st));
  count_launch(h, 2);
  n_g = d2h_scalar(h, dcount);
  FOFB_REQUIRE(n_g > 0, "no galaxy passes the redshift cut");
  const int ocols = 8;
  if (h->upload_pending) {  // columns 3.. of a split upload arrive on the side stream; nothing before here read them
    FOFB_CUDA(cudaStreamWaitEvent(st, h->upload_done, 0));
    h->upload_pending = false;
  }
  double* G = g_rows.alloc((size_t)n_g * ocols);
  FOFB_PROFILED(h, "k_build_rows", k_build_rows<<<grid_for((int64_t)n_g * ocols, B), B, 0, st>>>(n_g, gsrc, raw, s0.N_row.p, G, ocols));
  // candidate centres in priority order
  const int m0 = (int)s0.n_candidates;
  m_c = 0;
  int* csrc = c_src.alloc(std::max(m0, 1));
  if (m0 > 0) {
    int* cflag = flags.alloc(std::max(n, m0));
    k_cand_flag<<<grid_for(m0, B), B, 0, st>>>(m0, s0.cand_rows.p, cat.z_row.p, czmin, zmax_cen, dist ? own_hi_exclusive : INFINITY, cflag);
    FOFB_LAUNCH_CHECK(h);
    FOFB_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, s0.cand_rows.p, cflag, csrc, dcount, m0, st));
    FOFB_CUDA(cub::DeviceSelect::Flagged(cub_tmp(h, bytes), bytes, s0.cand_rows.p, cflag, csrc, dcount, m0, st));
    count_launch(h, 2);
    m_c = d2h_scalar(h, dcount);
  }
  double* C = c_rows.alloc((size_t)std::max(m_c, 1) * ocols);
  if (m_c > 0) {
    FOFB_PROFILED(h, "k_build_rows", k_build_rows<<<grid_for((int64_t)m_c * ocols, B), B, 0, st>>>(m_c, csrc, raw, s0.N_row.p, C, ocols));
  }
  // FoF() rows == stage-0 rows (nothing cut): the z-sorted catalogue and the theta_0.5 cell list are shared
  if (n_g == n) fof.adopt(&cat, &cells); else fof.adopt(nullptr, nullptr);
  Gv.data = G; Gv.rows = n_g; Gv.cols = ocols; Gv.rs = ocols; Gv.cs = 1;
  Cv.data = C; Cv.rows = m_c; Cv.cols = ocols; Cv.rs = ocols; Cv.cs = 1;
}

void Pipeline::finish(fofb_handle* h, const double* u, int mem, unsigned* mt_key, int* mt_pos) {
  cudaStream_t st = h->stream;
  const int B = 256;
  if (!began) throw Error(FOFB_ERR_STATE, "fofb_pipeline_finish called before fofb_pipeline_begin");
  const fofb_params& p = params;
  const bool trace = getenv("FOFB_TRACE") != nullptr;
  timespec t_a, t_b;
  if (trace) { cudaStreamSynchronize(st); clock_gettime(CLOCK_MONOTONIC, &t_a); }
  if (mt_key) fof.finish_mt(h, mt_key, mt_pos);
  else fof.finish_unit(h, u, mem);
  if (trace) {
    cudaStreamSynchronize(st);
    clock_gettime(CLOCK_MONOTONIC, &t_b);
    fprintf(stderr, "[trace] stage 4 %.2f ms\n", (t_b.tv_sec - t_a.tv_sec) * 1e3 + (t_b.tv_nsec - t_a.tv_nsec) * 1e-6);
    t_a = t_b;
  }
  K_out = 0;
  total_out = 0;
  finished = true;
  const ClusterList& F = fof.final_;
  const int K = F.K;
  if (K == 0) return;
  // stage 5 on the final clusters
  const int cols = 8;
  double* M = m_rows.alloc((size_t)F.total * cols);
  FOFB_PROFILED(h, "k_gather_members", k_gather_members<<<grid_for(F.total * cols, B), B, 0, st>>>(F.total, F.rows.p, fof.G.data, cols, M));
  double* cen = centres.alloc((size_t)K * 3);
  k_cluster_centres<<<grid_for(K, B), B, 0, st>>>(K, F.centre.p, fof.c_ra.p, fof.c_dec.p, fof.c_z.p, cen);
  FOFB_LAUNCH_CHECK(h);
  View Mv;
  Mv.data = M; Mv.rows = F.total; Mv.cols = cols; Mv.rs = cols; Mv.cs = 1;
  clean.three_sigma_core(h, p, K, F.off.p, F.total, Mv, cen, p.n_bins);
  // fof.py:514: keep clusters whose cleaned richness is still >= richness
  int* flag = flags.alloc(std::max(K, 1));
  int* iota = iotas.alloc(std::max(K, 1));
  int* dcount = counter.alloc(2);
  k_rich_flag<<<grid_for(K, B), B, 0, st>>>(K, clean.d_out_off.p, p.richness, flag, iota);
  FOFB_LAUNCH_CHECK(h);
  int* sel = selected.alloc(K);
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, iota, flag, sel, dcount, K, st));
  FOFB_CUDA(cub::DeviceSelect::Flagged(cub_tmp(h, bytes), bytes, iota, flag, sel, dcount, K, st));
  count_launch(h, 2);
  const int K2 = d2h_scalar(h, dcount);
  K_out = K2;
  if (K2 == 0) return;
  long long* len = lens.alloc((size_t)K2 + 1);
  long long* ooff = out_off.alloc((size_t)K2 + 1);
  k_sel_sizes<<<grid_for(K2, B), B, 0, st>>>(K2, sel, clean.d_out_off.p, len);
  FOFB_LAUNCH_CHECK(h);
  total_out = scan_offsets(h, len, ooff, K2);
  int* grow = out_grow.alloc((size_t)total_out);
  k_cleaned_rows<<<grid_for((int64_t)K2 * 32, 128), 128, 0, st>>>(K2, sel, F.off.p, F.rows.p, clean.d_out_off.p,
                                                                  clean.d_out_index.p, ooff, grow);
  FOFB_LAUNCH_CHECK(h);
  // stage 6 on the cleaned members
  double* M2 = m_rows2.alloc((size_t)total_out * cols);
  FOFB_PROFILED(h, "k_gather_members", k_gather_members<<<grid_for(total_out * cols, B), B, 0, st>>>(total_out, grow, fof.G.data, cols, M2));
  View M2v;
  M2v.data = M2; M2v.rows = total_out; M2v.cols = cols; M2v.rs = cols; M2v.cs = 1;
  int* dmax = counter.p + 1;
  FOFB_CUDA(cudaMemsetAsync(dmax, 0, sizeof(int), st));
  k_max_len<<<grid_for(K2, B), B, 0, st>>>(K2, ooff, dmax);
  FOFB_LAUNCH_CHECK(h);
  const int nmax = d2h_scalar(h, dmax);
  clean.virial_core(h, p, K2, ooff, M2v, nmax, props.alloc((size_t)K2 * 8));
  // results in terms of the caller's rows
  k_final_meta<<<grid_for(K2, B), B, 0, st>>>(K2, sel, F.centre.p, F.N.p, F.D.p, dist ? nullptr : c_src.p, g_src.p, fof.m,
                                              out_centre.alloc(K2), out_N.alloc(K2), out_D.alloc(K2));
  FOFB_LAUNCH_CHECK(h);
  k_map_rows<<<grid_for(total_out, B), B, 0, st>>>(total_out, grow, g_src.p, out_rows.alloc((size_t)total_out));
  FOFB_LAUNCH_CHECK(h);
  if (trace) {
    cudaStreamSynchronize(st);
    clock_gettime(CLOCK_MONOTONIC, &t_b);
    fprintf(stderr, "[trace] stages 5-6 %.2f ms\n", (t_b.tv_sec - t_a.tv_sec) * 1e3 + (t_b.tv_nsec - t_a.tv_nsec) * 1e-6);
  }
}

}  // namespace fofb
