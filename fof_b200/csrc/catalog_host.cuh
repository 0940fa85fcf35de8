// Owning side of the device catalogue (see catalog.cuh for the device views).
#pragma once
#include "catalog.cuh"

namespace fofb {

struct Catalog {
  int64_t n = 0;
  View rows;
  BBox gbox{};             // global (ra_min, ra_max, dec_min, dec_max)  (helper.py:197-202)
  double zmin = 0, zmax = 0;
  double cosmin = 1.0;
  int nblocks = 0;
  bool have_cells = false;
  double cell_radius = 0.0;  // the query radius the cell list was sized for
  double cell_size_ra = 1.0, cell_size_dec = 1.0;
  int ncx = 1, ncy = 1;

  DBuf<double> ra_row, dec_row, z_row;  // row order
  DBuf<double> zs, zra, zdec;           // z-rank order
  DBuf<int> zorder, rank, idx_tmp;
  DBuf<double4> pre, suf, levels, gbox_dev;
  DBuf<unsigned> key_a, key_b;
  DBuf<int> val_b;
  DBuf<double> c_ra, c_dec, c_cos;      // cell order
  DBuf<int> c_zrank, c_row, cell_start;

  void build(fofb_handle* h, const View& rows);         // z sort + range-bbox tables
  void build_cells(fofb_handle* h, double r_max_deg);   // (dec, ra) cell list sized for r_max
  CatalogView view() const;
};

}  // namespace fofb
