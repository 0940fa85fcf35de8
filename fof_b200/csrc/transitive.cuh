// State of the transitive (connected-components) mode, see transitive.cu.
#pragma once
#include "catalog_host.cuh"
#include "stage0.cuh"

namespace fofb {

struct Transitive {
  Catalog cat;
  CellList cells;
  DBuf<QueryRec> rec;
  DBuf<double> theta, theta_max;
  DBuf<int> parent, label, is_root, num_groups, run_start, run_head, run_end;
  bool used_runs = false;
  DBuf<unsigned long long> links;
  long long n_rows = 0, n_groups = 0, n_links = 0;
  void run(fofb_handle* h, const fofb_params& p, const View& rows, double ll_mpc_h, bool count_links);
};

}  // namespace fofb
