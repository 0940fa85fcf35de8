// Device-resident galaxy catalogue: z-sorted order, exact velocity windows, exact range
// min/max of (ra, dec) over windows, and the (dec, ra) cell list with z-sorted cells.
//
// Replaces, for every stage, the reference's per-centre "mask the whole catalogue, then build a
// GriSPy grid on the survivors" (fof.py:67-69,158-162; helper.py:181,207-211).
#pragma once
#include "common.cuh"
#include "geom.cuh"

namespace fofb {

constexpr int kBBoxBlock = 256;
#ifndef FOFB_ZBINS
#define FOFB_ZBINS 64
#endif
constexpr int kZBins = FOFB_ZBINS;       // z-rank bins per cell of a packed cell list (128 bytes of u16 offsets per cell)  // block size of the two-level range min/max structure

// One cell-list entry for the tiled bubble counts (16 bytes): offsets from the entry's own cell corner in degrees and
// cos(dec) in fp32, plus the z-rank.  Differences between entries of neighbouring cells are then exact to ~1e-8 deg in
// fp32 (the cell shift is added as a small multiple of the cell size), which is enough to decide every pair that is not
// within 1e-4 of the threshold; those few are re-decided in fp64 from the full-precision arrays.
struct PackedPoint {
  float fx, fy, fcos;
  int zr;
};

struct CellListView {
  double ra0, dec0, inv_sra, inv_sdec;
  double s_ra, s_dec;      // cell size in degrees
  int ncx, ncy;
  const PackedPoint* pk;   // cell-ordered, or null when the list was built without them
  // z-rank bins per cell (with pk): ztab[cell * kZBins + b] = offset inside the cell of its first entry whose
  // z-rank >> zshift is >= b.  A velocity window is about one bin wide, so the start of a cell's in-window run is one
  // table look-up and a step or two instead of a binary search.  Null when a cell holds 65535 entries or more.
  const unsigned short* ztab;
  int zshift;
  const int* cell_start;   // [ncx*ncy + 1]
  const double* ra;        // cell-ordered SoA; z-rank ascending inside a cell
  const double* dec;
  const double* cosdec;
  const int* zrank;
  const int* row;
  __device__ __forceinline__ int cx_of(double ra_) const {
    double t = floor((ra_ - ra0) * inv_sra);
    return t < 0.0 ? 0 : (t > (double)(ncx - 1) ? ncx - 1 : (int)t);
  }
  __device__ __forceinline__ int cy_of(double dec_) const {
    double t = floor((dec_ - dec0) * inv_sdec);
    return t < 0.0 ? 0 : (t > (double)(ncy - 1) ? ncy - 1 : (int)t);
  }
};

// conservative half-width in RA (degrees) of a bubble of radius r around declination dec
__host__ __device__ __forceinline__ double ra_reach(double r, double dec) {
  const double d = fmin(89.0, fabs(dec) + r);
  return 1.01 * r / cos(d * kDeg2Rad);
}

struct CatalogView {
  int64_t n;
  const double* zs;      // z ascending
  const double* zra;     // ra, dec in z-rank order
  const double* zdec;
  const int* zorder;     // row index of each z-rank
  const int* rank;       // z-rank of each row
  const double4* pre;    // per rank: (ra_min, ra_max, dec_min, dec_max) of block start..rank
  const double4* suf;    // per rank: rank..block end
  const double4* levels; // sparse table over blocks, level k at levels + k*nblocks
  int nblocks;
  double dec_absmax;     // largest |dec| in the catalogue: bounds the RA reach of a bubble
};

// Velocity window of a centre at redshift zc as the half-open z-rank range [lo, hi):
// {j : |redshift_to_velocity(z_j, zc)| <= vmax}, exact w.r.t. the reference predicate
// (fof.py:68,159,271; helper.py:208).  The velocity is monotone in z_j, so two binary searches do it.
__device__ __forceinline__ void velocity_window(const double* zs, int n, double zc, double vmax, int* lo, int* hi) {
  int a = 0, b = n;
  while (a < b) {  // first j with v >= -vmax
    const int m = (a + b) >> 1;
    if (redshift_to_velocity(zs[m], zc) < -vmax) a = m + 1; else b = m;
  }
  *lo = a;
  b = n;
  while (a < b) {  // first j with v > vmax
    const int m = (a + b) >> 1;
    if (redshift_to_velocity(zs[m], zc) <= vmax) a = m + 1; else b = m;
  }
  *hi = a;
}

__device__ __forceinline__ double4 bbox_merge(double4 a, double4 b) {
  return make_double4(fmin(a.x, b.x), fmax(a.y, b.y), fmin(a.z, b.z), fmax(a.w, b.w));
}

// exact (ra_min, ra_max, dec_min, dec_max) over z-ranks [lo, hi), hi > lo
__device__ __forceinline__ BBox range_bbox(const CatalogView& c, int lo, int hi) {
  const int bl = lo / kBBoxBlock, bh = (hi - 1) / kBBoxBlock;
  double4 v;
  if (bl == bh) {
    // same block: prefix/suffix cover the range only when it touches a block edge; else scan
    const int base = bl * kBBoxBlock;
    if (lo == base) {
      v = c.pre[hi - 1];
    } else if (hi == base + kBBoxBlock || hi == (int)c.n) {
      v = c.suf[lo];
    } else {
      v = make_double4(INFINITY, -INFINITY, INFINITY, -INFINITY);
      for (int j = lo; j < hi; ++j) {
        const double ra = c.zra[j], dec = c.zdec[j];
        v = bbox_merge(v, make_double4(ra, ra, dec, dec));
      }
    }
  } else {
    v = bbox_merge(c.suf[lo], c.pre[hi - 1]);
    const int a = bl + 1, b = bh;  // full blocks [a, b)
    if (b > a) {
      const int k = 31 - __clz(b - a);
      v = bbox_merge(v, bbox_merge(c.levels[(size_t)k * c.nblocks + a], c.levels[(size_t)k * c.nblocks + b - (1 << k)]));
    }
  }
  return BBox{v.x, v.y, v.z, v.w};
}

}  // namespace fofb
