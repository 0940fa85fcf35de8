// Device-resident whole-job pipeline state (see pipeline.cu).
#pragma once
#include "catalog_host.cuh"
#include "clean.cuh"
#include "fof_stage.cuh"
#include "stage0.cuh"

namespace fofb {

struct Pipeline {
  bool began = false, finished = false;
  fofb_params params{};
  Catalog cat;       // on the raw input (stage 0)
  CellList cells;
  Stage0 s0;
  FofState fof;      // on the z-cut main_arr
  CleanState clean;
  int n_in = 0, n_g = 0, m_c = 0;
  int K_out = 0;
  long long total_out = 0;
  DBuf<int> flags, iotas, counter, g_src, c_src, selected, out_grow, out_rows, out_centre;
  DBuf<double> g_rows, c_rows, m_rows, m_rows2, centres, props, out_N, out_D;
  DBuf<long long> lens, out_off;
  void begin(fofb_handle* h, const fofb_params& p, const View& raw, double zmin, double zmax_gal, double zmax_cen,
             const unsigned* mt_key = nullptr, int mt_pos = 0);
  void finish(fofb_handle* h, const double* u, int mem, unsigned* mt_key = nullptr, int* mt_pos = nullptr);
  void rerun(fofb_handle* h, double linking_length_factor, int richness, double overdensity, const unsigned* mt_key, int mt_pos);
  void export_local_clusters(fofb_handle* h, int* centre, long long* sizes, double* ids);
  void begin_dist(fofb_handle* h, const fofb_params& p, const View& raw, double zmin, double zmax_gal, double zmax_cen,
                  double valid_lo, double valid_hi, double own_lo, double own_hi);
  void stage0_and_inputs(fofb_handle* h, const fofb_params& p, const View& raw, double zmin, double zmax_gal, double czmin,
                         double zmax_cen, const unsigned* mt_key, int mt_pos);
  bool dist = false;
  double own_hi_exclusive = 0;
  View Gv, Cv;
  DBuf<double> cn_glob;
  DBuf<unsigned char> owned;
  DBuf<int> keep_dev, ext_centre;
  DBuf<long long> ext_off;
  DBuf<double> ext_ids;
};

}  // namespace fofb
