// Shared-memory tile staging with Blackwell bulk asynchronous copies (cp.async.bulk + mbarrier; SASS: UBLKCP / SYNCS).
// Used by the cell-tiled bubble-count kernels (counts.cu: N(0.5); overlap.cu: overdensity sample points).
#pragma once
#include <cuda_runtime.h>

#include "catalog.cuh"

namespace fofb {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// global -> shared bulk copy; src, dst 16-byte aligned, bytes a multiple of 16; completes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  }
}

template <class ZR>
__device__ __forceinline__ int lower_bound_zr(const ZR* __restrict__ a, int lo, int hi, int key) {
  while (lo < hi) {
    const int m = (lo + hi) >> 1;
    if (a[m] < key) lo = m + 1; else hi = m;
  }
  return lo;
}

// ---- one strip of cells and its neighbourhood ---------------------------------------------------------------------
// A strip is kCtW consecutive cells of one cell row.  Its neighbourhood -- the same cells plus one on either side, in
// the rows above, at and below -- is three contiguous slices of the cell-ordered packed records (PackedPoint, 16 bytes).
// Tile geometry (overridable for sweeps: scripts/build_tile_variants.py).  Measured at 10 M galaxies (53 per cell, 212
// per strip, 960 staged points on average, 1.3 % of the strips above 1 536): 128 threads and 28 KB per CTA keep seven
// CTAs resident per SM -- k_count 2.99 ms against 3.80 ms with 256 threads, 3 072 points and four CTAs; 1 024 points
// send a third of the strips down the global-memory path (3.41 ms), eight CTAs at 64 registers spill (3.50 ms).
#ifndef FOFB_CT_W
#define FOFB_CT_W 4
#endif
#ifndef FOFB_CT_THREADS
#define FOFB_CT_THREADS 128
#endif
#ifndef FOFB_CT_MAXPTS
#define FOFB_CT_MAXPTS 1536
#endif
#ifndef FOFB_CT_RESIDENT
#define FOFB_CT_RESIDENT 7
#endif
constexpr int kCtW = FOFB_CT_W, kCtThreads = FOFB_CT_THREADS, kCtMaxPts = FOFB_CT_MAXPTS, kCtResident = FOFB_CT_RESIDENT;
constexpr size_t kCtZtabBytes = (size_t)3 * (kCtW + 2) * kZBins * sizeof(unsigned short);
constexpr size_t kCtSmemBytes = (size_t)kCtMaxPts * sizeof(PackedPoint) + kCtZtabBytes;
constexpr float kHalfDeg2RadF = 0.00872664625997164788f;  // pi / 360

struct Tile {
  const PackedPoint* pk;   // the staged slices (shared memory), or the cell list's own array when they did not fit
  const unsigned short* ztab;  // z-rank bin offsets of the neighbourhood's cells: [row][cell - xa][kZBins], or null
  int zstride;             // cells per row in ztab (kCtW + 2 when staged; unused when ztab is the global table)
  bool ztab_staged;
  int (*cs)[kCtW + 3];     // global cell boundaries of the three rows, from cell xa on
  int* delta;              // index of an entry in pk = its global index + delta[row]
  int xa, ncs;
};

// Fills t.cs, then brings the three slices into `smem` (thread 0 issues the bulk copies, everybody waits on the
// mbarrier).  When the slices exceed kCtMaxPts entries (percolating blobs) nothing is copied: t.pk is the cell list's
// array and delta = 0.  Must be called by every thread of the CTA; bar and fits_sh are shared.
__device__ __forceinline__ void tile_stage(const CellListView& cl, int y, int x0, Tile& t, PackedPoint* smem,
                                           unsigned short* smem_ztab, unsigned long long* bar, int* fits_sh) {
  const int tid = threadIdx.x;
  t.xa = x0 > 0 ? x0 - 1 : 0;
  const int xb = x0 + kCtW < cl.ncx - 1 ? x0 + kCtW : cl.ncx - 1;  // staged cells xa..xb
  t.ncs = xb - t.xa + 2;                                          // their boundaries
  if (tid < 3 * (kCtW + 3)) {
    const int r = tid / (kCtW + 3), k = tid % (kCtW + 3);
    const int yy = y + r - 1;
    t.cs[r][k] = (yy >= 0 && yy < cl.ncy && k < t.ncs) ? __ldg(cl.cell_start + (size_t)yy * cl.ncx + t.xa + k) : 0;
  }
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  if (tid == 0) {
    int sbase = 0;
    int ga[3], la[3];
    for (int r = 0; r < 3; ++r) {
      ga[r] = t.cs[r][0];
      la[r] = t.cs[r][t.ncs - 1] - ga[r];
      t.delta[r] = sbase - ga[r];
      sbase += la[r];
    }
    const int fits = sbase <= kCtMaxPts ? 1 : 0;
    if (fits) {
      // the bin tables of the staged cells ride along: (ncs - 1) cells x 128 bytes per valid row
      const int ncell_row = t.ncs - 1;
      int zrows = 0;
      if (cl.ztab)
        for (int r = 0; r < 3; ++r) zrows += (y + r - 1 >= 0 && y + r - 1 < cl.ncy) ? 1 : 0;
      mbar_expect_tx(bar, (unsigned)sbase * (unsigned)sizeof(PackedPoint) + (unsigned)(zrows * ncell_row * kZBins * 2));
      for (int r = 0; r < 3; ++r)
        if (la[r]) bulk_g2s(smem + t.delta[r] + ga[r], cl.pk + ga[r], (unsigned)la[r] * (unsigned)sizeof(PackedPoint), bar);
      if (cl.ztab)
        for (int r = 0; r < 3; ++r) {
          const int yy = y + r - 1;
          if (yy < 0 || yy >= cl.ncy) continue;
          bulk_g2s(smem_ztab + (size_t)r * (kCtW + 2) * kZBins, cl.ztab + ((size_t)yy * cl.ncx + t.xa) * kZBins,
                   (unsigned)(ncell_row * kZBins * 2), bar);
        }
    } else {
      t.delta[0] = t.delta[1] = t.delta[2] = 0;
    }
    *fits_sh = fits;
  }
  __syncthreads();
  t.zstride = kCtW + 2;
  if (*fits_sh) {
    mbar_wait(bar, 0);
    t.pk = smem;
    t.ztab = cl.ztab ? smem_ztab : nullptr;
    t.ztab_staged = true;
  } else {
    t.pk = cl.pk;
    t.ztab = cl.ztab;
    t.ztab_staged = false;
  }
}

// In-window run [first, last) of one cell of the neighbourhood (row r, cell cx, entries [s, e) in t.pk; entries are
// z-rank sorted inside a cell).  With the bin table both ends are a look-up and a short walk (a walk that gets long -- a
// blob whose members share one bin -- turns into a binary search); without it, two binary searches.
__device__ __forceinline__ void tile_run(const CellListView& cl, const Tile& t, int y, int r, int cx, int s, int e, int wlo, int whi,
                                         int* first_out, int* last_out) {
  int first, last;
  if (t.ztab) {
    const unsigned short* tb = t.ztab_staged ? t.ztab + ((size_t)r * t.zstride + (cx - t.xa)) * kZBins
                                             : t.ztab + ((size_t)(y - 1 + r) * cl.ncx + cx) * kZBins;
    first = s + tb[wlo >> cl.zshift];
    const int hb = ((whi - 1) >> cl.zshift) + 1;
    last = hb < kZBins ? s + tb[hb] : e;
    for (int steps = 0; first < last && t.pk[first].zr < wlo; ++first)
      if (++steps > 12) {
        int lo = first, hi = last;
        while (lo < hi) {
          const int m = (lo + hi) >> 1;
          if (t.pk[m].zr < wlo) lo = m + 1; else hi = m;
        }
        first = lo;
        break;
      }
    for (int steps = 0; last > first && t.pk[last - 1].zr >= whi; --last)
      if (++steps > 12) {
        int lo = first, hi = last;
        while (lo < hi) {
          const int m = (lo + hi) >> 1;
          if (t.pk[m].zr < whi) lo = m + 1; else hi = m;
        }
        last = lo;
        break;
      }
  } else {
    int lo = s, hi = e;
    while (lo < hi) {
      const int m = (lo + hi) >> 1;
      if (t.pk[m].zr < wlo) lo = m + 1; else hi = m;
    }
    first = lo;
    hi = e;
    while (lo < hi) {
      const int m = (lo + hi) >> 1;
      if (t.pk[m].zr < whi) lo = m + 1; else hi = m;
    }
    last = lo;
  }
  *first_out = first;
  *last_out = last;
}

// One bubble query against the neighbourhood: the number of entries of the window [wlo, whi) within the bubble.
//   (cx, cy), q : the query point's cell and its packed offsets (q.fx / q.fy relative to that cell's corner);
//   exact(ra, dec, cos) : the same bubble in fp64 -- decides the pairs the fp32 test leaves open (|a - T| <= band * T);
//                 a functor, so that a caller can fetch what the fp64 test needs only when a pair asks for it;
//   bx          : the GriSPy cell box of the query when `boxed` (rare): those queries run in fp64 throughout;
//   rows r0..r1 of the neighbourhood (0 = row y - 1), cells cx0..cx1.
struct TileQuery {
  float fx, fy, fcos, T_lo, T_hi;
  int cx, cy, wlo, whi;
};

template <class Exact>
__device__ __forceinline__ int tile_count(const CellListView& cl, const Tile& t, int y, const TileQuery& q, int r0, int r1, int cx0,
                                          int cx1, Exact exact, bool boxed, const Box& bx) {
  int count = 0;
  const float sx = (float)cl.s_ra, sy = (float)cl.s_dec;
  for (int r = r0; r <= r1; ++r) {
    const int d = t.delta[r];
    const float oy = (float)(y - 1 + r - q.cy) * sy - q.fy;
    int s = t.cs[r][cx0 - t.xa] + d;
    for (int cx = cx0; cx <= cx1; ++cx) {
      const int e = t.cs[r][cx - t.xa + 1] + d;
      int first, last;
      tile_run(cl, t, y, r, cx, s, e, q.wlo, q.whi, &first, &last);
      const float ox = (float)(cx - q.cx) * sx - q.fx;
      for (int j = first; j < last; ++j) {
        const PackedPoint p = t.pk[j];
        bool hit;
        if (!boxed) {
          const float hx = kHalfDeg2RadF * (p.fx + ox), hy = kHalfDeg2RadF * (p.fy + oy);
          const float sl = hx * (1.0f - hx * hx * (1.0f / 6.0f)), sd = hy * (1.0f - hy * hy * (1.0f / 6.0f));
          const float a = sd * sd + (q.fcos * p.fcos) * (sl * sl);
          if (a < q.T_lo) hit = true;
          else if (a > q.T_hi) hit = false;
          else hit = exact(cl.ra[j - d], cl.dec[j - d], cl.cosdec[j - d]);
        } else {
          const double pra = cl.ra[j - d], pdec = cl.dec[j - d];
          hit = bx.contains(pra, pdec) && exact(pra, pdec, cl.cosdec[j - d]);
        }
        count += hit ? 1 : 0;
      }
      s = e;
    }
  }
  return count;
}

// fp32 thresholds of a bubble of radius r_deg: pairs with a = sin^2(d/2) outside [T_lo, T_hi] are decided in fp32.
// The band covers the rounding of the packed offsets (a few 1e-8 deg against r_deg) with two orders of magnitude to spare.
__device__ __forceinline__ void tile_thresholds(double r_deg, double cell_deg, float* T_lo, float* T_hi) {
  const double sh = sin(0.5 * r_deg * kDeg2Rad);
  const double T = sh * sh;
  const double band = fmax(1.0e-4, 1.0e-5 * cell_deg / r_deg);
  *T_lo = (float)(T * (1.0 - band));
  *T_hi = (float)(T * (1.0 + band));
}

}  // namespace fofb
