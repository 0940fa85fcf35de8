// Owning side of the device catalogue (see catalog.cuh for the device views).
#pragma once
#include "catalog.cuh"

namespace fofb {

struct Catalog;

// (dec, ra) cell list over the catalogue, cells z-rank sorted; sized for one query radius.
struct CellList {
  bool built = false;
  double radius = 0.0;  // the query radius the list was sized for
  double ra0 = 0, dec0 = 0, s_ra = 1, s_dec = 1;
  int ncx = 1, ncy = 1;
  DBuf<double> ra, dec, cosdec;
  DBuf<PackedPoint> pk;   // only when built with packed = true (the tiled count kernels)
  DBuf<unsigned short> ztab;
  DBuf<int> ztab_flag;
  bool has_ztab = false;
  int zshift = 0;
  DBuf<int> zrank, row, cell_start, val_a, val_b;
  DBuf<unsigned> key_a, key_b;
  void build(fofb_handle* h, const Catalog& cat, double r_max_deg, bool packed = false);
  bool has_packed = false;
  CellListView view() const;
};

struct Catalog {
  int64_t n = 0;
  View rows;
  BBox gbox{};             // global (ra_min, ra_max, dec_min, dec_max)  (helper.py:197-202)
  double zmin = 0, zmax = 0;
  double dec_absmax = 0;
  int nblocks = 0;

  DBuf<double> ra_row, dec_row, z_row;  // row order
  DBuf<double> zs, zra, zdec;           // z-rank order
  DBuf<int> zorder, rank, idx_tmp, sorted_flag;
  DBuf<double4> pre, suf, levels, gbox_dev;

  void build(fofb_handle* h, const View& rows);  // z sort + range-bbox tables
  CatalogView view() const;
};

}  // namespace fofb
