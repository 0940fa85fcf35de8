// Helper-level bubble counts (see bubble.cu).
#pragma once
#include "catalog.cuh"

namespace fofb {

struct Bubble {
  DBuf<unsigned char> mask;
  DBuf<double4> partial;
  DBuf<double> cen;
  DBuf<int> counts;
  void count(fofb_handle* h, const fofb_params& p, const View& P, int q, const double* centres_host, const double* radius_host,
             int use_window, double zc, int* counts_host);
};

}  // namespace fofb
