// Catalogue matching (see match.cu).
#pragma once
#include "catalog.cuh"

namespace fofb {

struct Matcher {
  DBuf<double> z_in, z_sorted, x, y, z, zrange;
  DBuf<int> idx_in, row_sorted, match_row, counter;
  DBuf<double> match_sep;
  DBuf<unsigned char> cub_tmp;
  long long n_matched = 0;
  // sample: (S, >=3) [ra, dec, z]; catalogue: (C, >=3) [ra, dec, z].  Results stay on the device until copied out.
  void run(fofb_handle* h, const fofb_params& p, const View& sample, const View& catalogue, double distance_mpc_h, double z_cut);
};

}  // namespace fofb
