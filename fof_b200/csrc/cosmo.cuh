// Flat/curved LambdaCDM distances in fp64, usable on host and device.
//
// Replaces the astropy calls on the hot path (analysis/methods.py:31,59-64; fof.py:165-175,430;
// analysis/mass_estimator.py:86): angular_diameter_distance, comoving_transverse_distance,
// arcsec_per_kpc_proper and the quad() over differential_comoving_volume.
//
// D_C(z) = D_H * int_0^z dz'/E(z') is evaluated as a tabulated value at the panel edge below z
// (panels of width 1/64, table built once on the host with 16-point Gauss-Legendre per panel and
// compensated summation) plus one 8-point Gauss-Legendre rule over the remaining fraction of a
// panel.  The integrand is smooth, so both rules are exact to fp64 rounding (checked against
// scipy's hyp2f1 closed form in tests/test_cosmology.py to <= 4e-16 relative).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define FOFB_HD __host__ __device__ __forceinline__
#else
#define FOFB_HD inline
#endif

namespace fofb {

constexpr double kCkms = 299792.458;
constexpr double kPi = 3.14159265358979323846;
constexpr double kDeg2Rad = kPi / 180.0;   // numpy deg2rad multiplies by this constant
constexpr double kRad2Deg = 180.0 / kPi;   // numpy rad2deg multiplies by this constant
constexpr int kCosmoPanelsPerUnit = 64;
constexpr int kCosmoTableSize = 64 * 20 + 1;  // z up to 20

struct Cosmo {
  double H0, Om0, Ode0, Ok0, h, DH;
  double zmax;
  const double* table;  // D_C/D_H at panel edges (device or host pointer, matching the caller)
};

FOFB_HD double cosmo_inv_E(const Cosmo& c, double z) {
  const double zp1 = 1.0 + z;
  return 1.0 / sqrt(zp1 * zp1 * (c.Om0 * zp1 + c.Ok0) + c.Ode0);
}

// 8-point Gauss-Legendre on [a, b] of 1/E
FOFB_HD double cosmo_gl8_invE(const Cosmo& c, double a, double b) {
  const double x[4] = {0.1834346424956498049394761, 0.5255324099163289858177390,
                       0.7966664774136267395915539, 0.9602898564975362316835609};
  const double w[4] = {0.3626837833783619829651504, 0.3137066458778872873379622,
                       0.2223810344533744705443560, 0.1012285362903762591525314};
  const double hm = 0.5 * (b - a), mid = 0.5 * (a + b);
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s += w[i] * (cosmo_inv_E(c, mid - hm * x[i]) + cosmo_inv_E(c, mid + hm * x[i]));
  }
  return s * hm;
}

// line-of-sight comoving distance [Mpc]
FOFB_HD double cosmo_comoving_distance(const Cosmo& c, double z) {
  if (!(z > 0.0)) return z == 0.0 ? 0.0 : -cosmo_gl8_invE(c, z, 0.0) * c.DH;
  double zz = z < c.zmax ? z : c.zmax;
  int k = (int)(zz * kCosmoPanelsPerUnit);
  if (k > kCosmoTableSize - 1) k = kCosmoTableSize - 1;
  const double zk = (double)k / kCosmoPanelsPerUnit;
  double d = c.table[k] + cosmo_gl8_invE(c, zk, zz);
  if (z > c.zmax) {  // beyond the table: panels of growing width keep the rule accurate
    double a = c.zmax;
    while (a < z) {
      double b = a * 1.25 < z ? a * 1.25 : z;
      d += cosmo_gl8_invE(c, a, b);
      a = b;
    }
  }
  return d * c.DH;
}

// transverse comoving distance [Mpc]
FOFB_HD double cosmo_transverse_distance(const Cosmo& c, double z) {
  const double dc = cosmo_comoving_distance(c, z);
  if (c.Ok0 == 0.0) return dc;
  const double sk = sqrt(fabs(c.Ok0));
  const double x = sk * dc / c.DH;
  return c.Ok0 > 0.0 ? c.DH / sk * sinh(x) : c.DH / sk * sin(x);
}

FOFB_HD double cosmo_angular_diameter_distance(const Cosmo& c, double z) {
  return cosmo_transverse_distance(c, z) / (1.0 + z);
}

// analysis/methods.py:12-32 for a proper length in plain Mpc -> degrees
FOFB_HD double cosmo_proper_mpc_to_deg(const Cosmo& c, double d_mpc, double z) {
  return d_mpc / cosmo_angular_diameter_distance(c, z) * kRad2Deg;
}

// analysis/methods.py:12-32: proper h^-1 Mpc -> degrees
FOFB_HD double cosmo_linear_to_angular_dist(const Cosmo& c, double d_mpc_h, double z) {
  return cosmo_proper_mpc_to_deg(c, d_mpc_h / c.h, z);
}

// int_z^{z+dz} D_H * D_M(z')^2 / E(z') dz'   [Mpc^3 / sr]   (analysis/methods.py:59-63)
FOFB_HD double cosmo_shell_volume_per_sr(const Cosmo& c, double z, double dz) {
  const double a = cosmo_comoving_distance(c, z);
  const double inc = cosmo_gl8_invE(c, z, z + dz) * c.DH;  // D_C(z+dz) - D_C(z), no cancellation
  if (c.Ok0 == 0.0) {
    const double b = a + inc;
    return inc * (a * a + a * b + b * b) / 3.0;  // (b^3 - a^3) / 3
  }
  const double x[4] = {0.1834346424956498049394761, 0.5255324099163289858177390,
                       0.7966664774136267395915539, 0.9602898564975362316835609};
  const double w[4] = {0.3626837833783619829651504, 0.3137066458778872873379622,
                       0.2223810344533744705443560, 0.1012285362903762591525314};
  const double hm = 0.5 * dz, mid = z + hm;
  double s = 0.0;
  for (int i = 0; i < 4; ++i) {
    for (int sgn = -1; sgn <= 1; sgn += 2) {
      const double zz = mid + sgn * hm * x[i];
      const double dm = cosmo_transverse_distance(c, zz);
      s += w[i] * c.DH * dm * dm * cosmo_inv_E(c, zz);
    }
  }
  return s * hm;
}

// analysis/methods.py:35-65: mean inter-galaxy separation [Mpc]
FOFB_HD double cosmo_mean_separation(const Cosmo& c, double n, double z, double max_velocity, double survey_area) {
  const double dz = max_velocity / kCkms;
  const double ang = survey_area * kDeg2Rad;
  const double solid_angle = kPi * ang * ang;
  const double V = cosmo_shell_volume_per_sr(c, z, dz) * solid_angle;
  return 1.0 / cbrt(n / V);
}

// analysis/methods.py:85-87, with the reference's association and no FMA contraction, so the
// velocity-window predicate is bit-identical to numpy's on every platform.
FOFB_HD double redshift_to_velocity(double z, double center_z) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(__dsub_rn(__dmul_rn(kCkms, z), __dmul_rn(kCkms, center_z)), __dadd_rn(1.0, center_z));
#else
  volatile double a = kCkms * z;
  volatile double b = kCkms * center_z;
  volatile double num = a - b;
  volatile double den = 1.0 + center_z;
  return num / den;
#endif
}

// Host-side table builder: D_C/D_H at z_k = k/64 (16-point GL per panel, Neumaier summation).
inline void cosmo_build_table(double Om0, double Ode0, double* table /* kCosmoTableSize */) {
  static const double x16[8] = {0.0950125098376374401853193, 0.2816035507792589132304605,
                                0.4580167776572273863424194, 0.6178762444026437484466718,
                                0.7554044083550030338951012, 0.8656312023878317438804679,
                                0.9445750230732325760779884, 0.9894009349916499325961542};
  static const double w16[8] = {0.1894506104550684962853967, 0.1826034150449235888667637,
                                0.1691565193950025381893121, 0.1495959888165767320815017,
                                0.1246289712555338720524763, 0.0951585116824927848099251,
                                0.0622535239386478928628438, 0.0271524594117540948517806};
  const double Ok0 = 1.0 - Om0 - Ode0;
  auto invE = [&](long double z) -> long double {
    long double zp1 = 1.0L + z;
    return 1.0L / sqrtl(zp1 * zp1 * ((long double)Om0 * zp1 + (long double)Ok0) + (long double)Ode0);
  };
  long double acc = 0.0L;
  table[0] = 0.0;
  for (int k = 1; k < kCosmoTableSize; ++k) {
    const long double a = (long double)(k - 1) / kCosmoPanelsPerUnit, b = (long double)k / kCosmoPanelsPerUnit;
    const long double hm = 0.5L * (b - a), mid = 0.5L * (a + b);
    long double s = 0.0L;
    for (int i = 0; i < 8; ++i) s += (long double)w16[i] * (invE(mid - hm * x16[i]) + invE(mid + hm * x16[i]));
    acc += s * hm;
    table[k] = (double)acc;
  }
}

}  // namespace fofb
