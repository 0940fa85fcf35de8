// Stage 5/6 scratch kept in the handle (see clean.cu).
#pragma once
#include <vector>
#include "common.cuh"
#include "geom.cuh"

namespace fofb {

struct CleanState {
  DBuf<long long> d_off, d_len, d_out_off;
  DBuf<double> d_cen, d_vel, d_rad, d_dev, d_dA, d_props;
  DBuf<int> d_perm, d_bin_start, d_kept, d_out_index, d_status;
  DBuf<unsigned> d_mt;
  DBuf<int> d_big_count, d_big_begin, d_big_len, d_big_slot, d_big_active, d_big_seg, d_big_perm;
  DBuf<long long> d_big_off;
  DBuf<double> d_big_dev, d_big_psum;
  DBuf<double4> d_trig;
  DBuf<unsigned short> d_boot_table;  // bootstrap indices per cluster size (clean.cu: k_boot_table)
  DBuf<int> d_present;
  std::vector<char> boot_built;
  size_t mt_len = 0;
  size_t clip_smem_set = 0;
  // device-resident cores: results stay in d_out_off (K+1) / d_out_index; returns the kept total
  long long three_sigma_core(fofb_handle* h, const fofb_params& p, int K, const long long* d_off_in, long long total,
                             const View& M, const double* d_cen_in, int nbins);
  void clip_big_bins(fofb_handle* h, int K, int nbins, const long long* off, long long total, const double* vel,
                     const double* rad, double* dev, int* perm, const int* bstart, int* kept);
  void virial_core(fofb_handle* h, const fofb_params& p, int K, const long long* d_off_in, const View& M, int nmax,
                   double* d_out);
  void three_sigma(fofb_handle* h, const fofb_params& p, int K, const long long* off_host, const View& M,
                   const double* centres_host, int nbins, long long* out_off_host, int* out_index_host);
  void virial(fofb_handle* h, const fofb_params& p, int K, const long long* off_host, const View& M, double* out_host);
  void projected(fofb_handle* h, const fofb_params& p, int K, const long long* off_host, const View& M, double* out_host);
};

void bootstrap_indices_host(long long n, long long count, int* out);

}  // namespace fofb
