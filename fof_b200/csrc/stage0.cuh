// Stage 0 state (number counts + candidate ordering) kept in the handle between calls.
#pragma once
#include "common.cuh"

namespace fofb {

struct QueryRec {  // one bubble query: cell-box bounds, radius, z-rank window
  double ra_lo, ra_hi, dec_lo, dec_hi;
  double r_deg;
  int wlo, whi;
};

struct CountRec {  // stage-0 query of one galaxy (16 bytes): radius, z-rank window; bit 31 of whi: the cell box binds
  double r_deg;
  int wlo, whi;
};

struct Catalog;
struct CellList;

struct Stage0 {
  bool valid = false;
  int64_t n_rows = 0;
  int64_t n_candidates = 0;
  DBuf<CountRec> rec;
  DBuf<double4> box;   // cell-box bounds of the few queries whose box binds
  DBuf<double> theta, theta_max;
  DBuf<int> N_row, cand_rows, num_sel;
  DBuf<unsigned long long> key_all, key_sel, key_sorted;
  void run(fofb_handle* h, const fofb_params& p, Catalog& cat, CellList& cells);
};

}  // namespace fofb
