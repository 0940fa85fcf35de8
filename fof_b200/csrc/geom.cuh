// Spherical geometry and GriSPy grid predicates in fp64 (device side).
//
// Replaces GriSPy's haversine metric and cell-box candidate selection (call sites fof.py:177,203,
// 209,280,337; helper.py:183,213; algorithm restated in SURVEY.md App. B.1) and astropy's
// SkyCoord.separation (Vincenty; fof.py:429, analysis/mass_estimator.py:88).
#pragma once
#include "cosmo.cuh"

namespace fofb {

// ---- GriSPy haversine, literal association (degrees in, degrees out) ---------------------------
// numpy evaluates  sdlat**2 + (clat1*clat2)*(sdlon**2)  with separate roundings; the explicit _rn
// intrinsics keep nvcc from contracting them into FMAs.
__device__ __forceinline__ double haversine_a_literal(double ra1, double dec1, double cosdec1, double ra2,
                                                      double dec2, double cosdec2) {
  const double lon1 = __dmul_rn(ra1, kDeg2Rad), lat1 = __dmul_rn(dec1, kDeg2Rad);
  const double lon2 = __dmul_rn(ra2, kDeg2Rad), lat2 = __dmul_rn(dec2, kDeg2Rad);
  const double sdlon = sin(__dmul_rn(__dsub_rn(lon2, lon1), 0.5));
  const double sdlat = sin(__dmul_rn(__dsub_rn(lat2, lat1), 0.5));
  return __dadd_rn(__dmul_rn(sdlat, sdlat), __dmul_rn(__dmul_rn(cosdec1, cosdec2), __dmul_rn(sdlon, sdlon)));
}

__device__ __forceinline__ double haversine_deg_literal(double ra1, double dec1, double cosdec1, double ra2,
                                                        double dec2, double cosdec2) {
  const double a = haversine_a_literal(ra1, dec1, cosdec1, ra2, dec2, cosdec2);
  return __dmul_rn(__dmul_rn(2.0, asin(sqrt(a))), kRad2Deg);
}

// sin(x) for |x| << 1 by its Taylor series (rel. error < 1e-16 for |x| < 0.02), libm otherwise
__device__ __forceinline__ double sin_small(double x) {
  if (fabs(x) < 0.02) {
    const double x2 = x * x;
    return x * (1.0 + x2 * (-1.0 / 6.0 + x2 * (1.0 / 120.0 + x2 * (-1.0 / 5040.0 + x2 * (1.0 / 362880.0)))));
  }
  return sin(x);
}

// hav(c, p) <= r  decided on  a = sin^2(d/2)  against  T = sin^2(r/2); pairs closer to the threshold
// than the guard band are re-decided with the literal degrees formula (bit-for-bit GriSPy association).
struct BubbleTest {
  double ra, dec, cosdec, r_deg, T_lo, T_hi;
  __device__ __forceinline__ void init(double ra_, double dec_, double cosdec_, double r) {
    const double s = sin(0.5 * r * kDeg2Rad);
    init_T(ra_, dec_, cosdec_, r, s * s);
  }
  // with T = sin^2(r/2) precomputed (r in degrees)
  __device__ __forceinline__ void init_T(double ra_, double dec_, double cosdec_, double r, double T) {
    ra = ra_; dec = dec_; cosdec = cosdec_; r_deg = r;
    T_lo = T * (1.0 - 1e-9);
    T_hi = T * (1.0 + 1e-9);
  }
  __device__ __forceinline__ bool inside(double pra, double pdec, double pcos) const {
    const double sl = sin_small(0.5 * kDeg2Rad * (pra - ra));
    const double sd = sin_small(0.5 * kDeg2Rad * (pdec - dec));
    const double a = sd * sd + cosdec * pcos * sl * sl;
    if (a < T_lo) return true;
    if (a > T_hi) return false;
    return haversine_deg_literal(ra, dec, cosdec, pra, pdec, pcos) <= r_deg;
  }
};

// ---- astropy angular_separation (Vincenty), radians ---------------------------------------------
// num2 must be exactly 0 for identical coordinates (three_sigma's "is this the centre" test,
// fof.py:462), hence the explicit roundings.
__device__ __forceinline__ double vincenty_rad(double lon1, double lat1, double lon2, double lat2) {
  double sdlon, cdlon, slat1, clat1, slat2, clat2;
  sincos(__dsub_rn(lon2, lon1), &sdlon, &cdlon);
  sincos(lat1, &slat1, &clat1);
  sincos(lat2, &slat2, &clat2);
  const double num1 = __dmul_rn(clat2, sdlon);
  const double num2 = __dsub_rn(__dmul_rn(clat1, slat2), __dmul_rn(__dmul_rn(slat1, clat2), cdlon));
  const double den = __dadd_rn(__dmul_rn(slat1, slat2), __dmul_rn(__dmul_rn(clat1, clat2), cdlon));
  return atan2(hypot(num1, num2), den);
}

// ---- GriSPy grid: bins = linspace(min - 1e-6, max + 1e-6, N + 1) ----------------------------------
struct GridAxis {
  double b0, bN;  // bins[0], bins[-1]
  int ncell;
  __host__ __device__ __forceinline__ void init(double vmin, double vmax, int n) {
    b0 = vmin - 1.0e-6;
    bN = vmax + 1.0e-6;
    ncell = n;
  }
  // int(N * (x - b0) / (bN - b0)) with C truncation, as a double (no overflow for far-away x)
  __device__ __forceinline__ double digit(double x) const {
    return trunc(__ddiv_rn(__dmul_rn((double)ncell, __dsub_rn(x, b0)), __dsub_rn(bN, b0)));
  }
  __device__ __forceinline__ int digit_clamped(double x) const {
    double t = digit(x);
    t = t < 0.0 ? 0.0 : t;
    t = t > (double)(ncell - 1) ? (double)(ncell - 1) : t;
    return (int)t;
  }
};

// Closed interval [lo, hi] of coordinate values whose cell index lies in [ilo, ihi]
// (ilo = clamp(digit(c - r)), ihi = clamp(digit(c + r))).  digit() is monotone, so the cell-box
// test on a point is equivalent to lo <= x <= hi; the bounds are found by stepping single ulps
// around the analytic cell edge until the exact predicate flips.
//
// `reach` >= the largest |x - c| a point can have and still pass the haversine test the box is always combined
// with.  When c - reach falls in the same cell as c - r the lower side can never exclude such a point, so the
// bound is -inf and the (division-heavy) edge search is skipped; likewise above.  This is the common case: an edge
// must fall inside the sliver (r, reach] for the box to bind at all.
__device__ __forceinline__ void grid_axis_bounds(const GridAxis& g, double c, double r, double reach, double* lo, double* hi) {
  const double dlo = g.digit(__dsub_rn(c, r)), dhi = g.digit(__dadd_rn(c, r));
  const double top = (double)(g.ncell - 1);
  const int ilo = dlo < 0.0 ? 0 : (dlo > top ? g.ncell - 1 : (int)dlo);
  const int ihi = dhi < 0.0 ? 0 : (dhi > top ? g.ncell - 1 : (int)dhi);
  const double w = (g.bN - g.b0) / g.ncell;
  if (ilo <= 0 || (dlo <= top && g.digit(c - reach) == dlo)) {
    *lo = -INFINITY;  // every member of the grid's point set has digit >= 0
  } else {
    double x = g.b0 + ilo * w;
    for (int it = 0; it < 64 && g.digit(x) >= (double)ilo; ++it) x = nextafter(x, -INFINITY);
    for (int it = 0; it < 64 && g.digit(x) < (double)ilo; ++it) x = nextafter(x, INFINITY);
    *lo = x;
  }
  if (ihi >= g.ncell - 1 || (dhi >= 0.0 && g.digit(c + reach) == dhi)) {
    *hi = INFINITY;
  } else {
    double x = g.b0 + (ihi + 1) * w;
    for (int it = 0; it < 64 && g.digit(x) <= (double)ihi; ++it) x = nextafter(x, INFINITY);
    for (int it = 0; it < 64 && g.digit(x) > (double)ihi; ++it) x = nextafter(x, -INFINITY);
    *hi = x;
  }
}

// Shortcut for the common case of grid_axis_bounds: true when NEITHER side of the cell box can bind, decided on the
// approximate quotients t = (x - b0) * inv_w (inv_w = ncell / (bN - b0)) whenever they are farther than `rel` from the
// integers that separate the outcomes -- a few ulps of rounding cannot move them across.  False means "not decided
// here": the caller runs the exact routine.
__device__ __forceinline__ bool grid_axis_unbound_fast(const GridAxis& g, double inv_w, double c, double r, double reach,
                                                       double rel = 1.0e-9) {
  const double top = (double)(g.ncell - 1);
  const double t1 = (c - r - g.b0) * inv_w, t2 = (c - reach - g.b0) * inv_w;   // lower side: t2 <= t1
  const double t3 = (c + r - g.b0) * inv_w, t4 = (c + reach - g.b0) * inv_w;   // upper side: t4 >= t3
  auto clear = [rel](double t) {  // not within the margin of an integer
    const double f = t - floor(t), m = rel * (fabs(t) + 1.0);
    return f > m && f < 1.0 - m;
  };
  const double m1 = rel * (fabs(t1) + 1.0), m3 = rel * (fabs(t3) + 1.0);
  const bool lower_free = (t1 < 1.0 - m1) ||
                          (t2 > 0.0 && floor(t1) == floor(t2) && floor(t1) <= top && clear(t1) && clear(t2));
  const bool upper_free = (t3 > top + m3) || (t3 > m3 && floor(t3) == floor(t4) && clear(t3) && clear(t4));
  return lower_free && upper_free;
}

struct Box {  // cell-box bounds of one query in raw coordinates
  double ra_lo, ra_hi, dec_lo, dec_hi;
  __device__ __forceinline__ bool contains(double ra, double dec) const {
    return ra >= ra_lo && ra <= ra_hi && dec >= dec_lo && dec <= dec_hi;
  }
};

struct BBox {  // bounding box of a point set: (ra_min, ra_max, dec_min, dec_max)
  double ra_min, ra_max, dec_min, dec_max;
};

// The box of a bubble query (centre, r); only valid in conjunction with the haversine test d <= r.
__device__ __forceinline__ Box grid_box(const BBox& bb, int ncell, double c_ra, double c_dec, double r) {
  GridAxis gx, gy;
  gx.init(bb.ra_min, bb.ra_max, ncell);
  gy.init(bb.dec_min, bb.dec_max, ncell);
  Box b;
  // d <= r bounds |ddec| by r and |dra| by r / cos(|dec| + r) (1 % margin inside bubble_ra_reach)
  const double dd = fmin(89.0, fabs(c_dec) + r);
  const double reach_ra = 1.01 * r / cos(dd * kDeg2Rad) + 1e-12;
  grid_axis_bounds(gx, c_ra, r, reach_ra, &b.ra_lo, &b.ra_hi);
  grid_axis_bounds(gy, c_dec, r, r * (1.0 + 1e-9) + 1e-12, &b.dec_lo, &b.dec_hi);
  return b;
}

// The same with the shortcut first: inv_wx / inv_wy = ncell / (bN - b0) of the two axes, precomputed by the caller.
// *bound is false when the box is known not to bind (its four bounds are then infinite).
__device__ __forceinline__ Box grid_box_fast(const BBox& bb, int ncell, double inv_wx, double inv_wy, double c_ra, double c_dec,
                                             double r, bool* bound) {
  GridAxis gx, gy;
  gx.init(bb.ra_min, bb.ra_max, ncell);
  gy.init(bb.dec_min, bb.dec_max, ncell);
  const double dd = fmin(89.0, fabs(c_dec) + r);
  const double reach_ra = 1.01 * r / cos(dd * kDeg2Rad) + 1e-12;
  if (grid_axis_unbound_fast(gx, inv_wx, c_ra, r, reach_ra) && grid_axis_unbound_fast(gy, inv_wy, c_dec, r, r * (1.0 + 1e-9) + 1e-12)) {
    *bound = false;
    return Box{-INFINITY, INFINITY, -INFINITY, INFINITY};
  }
  const Box b = grid_box(bb, ncell, c_ra, c_dec, r);
  *bound = b.ra_lo != -INFINITY || b.ra_hi != INFINITY || b.dec_lo != -INFINITY || b.dec_hi != INFINITY;
  return b;
}

// GriSPy cell index of a point in a grid (for the cell-major return order)
__device__ __forceinline__ int grid_cell_key(const BBox& bb, int ncell, double ra, double dec) {
  GridAxis gx, gy;
  gx.init(bb.ra_min, bb.ra_max, ncell);
  gy.init(bb.dec_min, bb.dec_max, ncell);
  return gy.digit_clamped(dec) * ncell + gx.digit_clamped(ra);
}

}  // namespace fofb
