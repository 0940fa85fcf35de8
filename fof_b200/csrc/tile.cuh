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
// the rows above, at and below -- is three contiguous slices of the cell-ordered SoA (ra, dec, cos dec, z-rank).
constexpr int kCtW = 8, kCtThreads = 512, kCtMaxPts = 3584;
constexpr size_t kCtSmemBytes = (size_t)kCtMaxPts * (3 * sizeof(double) + sizeof(int));

struct Tile {
  double *ra, *dec, *cos;  // shared-memory slices (row r of the neighbourhood is shifted by delta[r])
  int* zr;
  int (*cs)[kCtW + 3];     // global cell boundaries of the three rows, from cell xa on
  int* delta;
  int xa, ncs;
};

// Fills t.cs, then brings the three slices into shared memory (thread 0 issues the bulk copies, everybody waits on the
// mbarrier).  Returns false -- with delta = 0, so that the caller indexes the cell list itself -- when the slices
// exceed kCtMaxPts points (percolating blobs).  Must be called by every thread of the CTA; bar and fits are shared.
__device__ __forceinline__ bool tile_stage(const CellListView& cl, int y, int x0, Tile& t, unsigned long long* bar, int* fits_sh) {
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
    int sbase = 0, fits = 1;
    int ga[3], la[3];
    for (int r = 0; r < 3; ++r) {
      const int gs = t.cs[r][0], ge = t.cs[r][t.ncs - 1];
      ga[r] = gs & ~3;  // 16-byte aligned slices: 4 ints, 4 doubles (the buffers carry slack past their last element)
      la[r] = ge > gs ? ((ge + 3) & ~3) - ga[r] : 0;
      t.delta[r] = sbase - ga[r];
      sbase += la[r];
    }
    if (sbase > kCtMaxPts) fits = 0;
    if (fits) {
      mbar_expect_tx(bar, (unsigned)sbase * 28u);
      for (int r = 0; r < 3; ++r) {
        if (!la[r]) continue;
        const int so = t.delta[r] + ga[r];
        bulk_g2s(t.ra + so, cl.ra + ga[r], (unsigned)la[r] * 8u, bar);
        bulk_g2s(t.dec + so, cl.dec + ga[r], (unsigned)la[r] * 8u, bar);
        bulk_g2s(t.cos + so, cl.cosdec + ga[r], (unsigned)la[r] * 8u, bar);
        bulk_g2s(t.zr + so, cl.zrank + ga[r], (unsigned)la[r] * 4u, bar);
      }
    } else {
      t.delta[0] = t.delta[1] = t.delta[2] = 0;
    }
    *fits_sh = fits;
  }
  __syncthreads();
  const bool fits = *fits_sh != 0;
  if (fits) mbar_wait(bar, 0);
  return fits;
}

// Number of points of the window [wlo, whi) inside the bubble `bt` (and the cell box `bx`) among the cells
// (rows r0..r1 of the neighbourhood, columns k0..k1 counted from t.xa).  P_* are the staged slices or the cell list.
__device__ __forceinline__ int tile_count(const double* __restrict__ P_ra, const double* __restrict__ P_dec,
                                          const double* __restrict__ P_cos, const int* __restrict__ P_zr, const Tile& t,
                                          int r0, int r1, int k0, int k1, int wlo, int whi, const BubbleTest& bt, const Box& bx) {
  int count = 0;
  for (int r = r0; r <= r1; ++r) {
    const int d = t.delta[r];
    int s = t.cs[r][k0] + d;
    for (int k = k0; k <= k1; ++k) {
      const int e = t.cs[r][k + 1] + d;
      const int a = lower_bound_zr(P_zr, s, e, wlo);
      const int b = lower_bound_zr(P_zr, a, e, whi);
      for (int j = a; j < b; ++j) {
        const double pra = P_ra[j], pdec = P_dec[j];
        if (bx.contains(pra, pdec) && bt.inside(pra, pdec, P_cos[j])) ++count;
      }
      s = e;
    }
  }
  return count;
}

}  // namespace fofb
