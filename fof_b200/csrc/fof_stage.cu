// Stage 1 of FoF() (fof.py:145-246): per-centre virial bubble, hop 1, hop 2 with the element-wise
// novelty rule, richness gate and the greedy tracker, evaluated as deterministic
// priority rounds: a centre is evaluated once every higher-priority centre that could claim it as a
// member has been decided (kernels K7-K9 of DESIGN.md).
#include <cub/cub.cuh>

#include <ctime>

#include "fof_stage.cuh"

namespace fofb {

static unsigned char* cub_tmp(fofb_handle* h, size_t bytes) { return h->cub_tmp.alloc(bytes + 256); }



// ---- per-centre scalars (fof.py:158-175, 277-279; helper.py:182) -------------------------------------
__global__ void k_centre_prep(View Cn, int m, CatalogView cat, Cosmo cosmo, double vmax, double virial_radius,
                              double count_radius, double* c_ra, double* c_dec, double* c_cos, double* c_z,
                              double* c_R1, double* c_T1, double* c_thvir, double* c_th05, double* c_mag, double* c_N,
                              double* c_id, int* c_wlo, int* c_whi, unsigned char* alive, unsigned char* decided) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double ra = Cn.at(i, 0), dec = Cn.at(i, 1), z = Cn.at(i, 2);
  c_ra[i] = ra;
  c_dec[i] = dec;
  c_cos[i] = cos(__dmul_rn(dec, kDeg2Rad));
  c_z[i] = z;
  c_mag[i] = Cn.at(i, 5);
  c_id[i] = Cn.at(i, 6);
  c_N[i] = Cn.at(i, Cn.cols - 1);  // cluster.py:38: N = center[-1]
  const double thvir = cosmo_linear_to_angular_dist(cosmo, virial_radius, z);
  // fof.py:165-175: angle -> comoving transverse length -> angle again (= theta_vir * (1 + z), Q4)
  const double ang = thvir * kDeg2Rad;
  const double max_dist_mpc = ang * cosmo_transverse_distance(cosmo, z);
  const double R1 = cosmo_proper_mpc_to_deg(cosmo, max_dist_mpc, z);
  c_R1[i] = R1;
  {
    const double sh = sin(0.5 * R1 * kDeg2Rad);
    c_T1[i] = sh * sh;
  }
  c_thvir[i] = thvir;
  c_th05[i] = cosmo_linear_to_angular_dist(cosmo, count_radius, z);
  int lo, hi;
  velocity_window(cat.zs, (int)cat.n, z, vmax, &lo, &hi);
  c_wlo[i] = lo;
  c_whi[i] = hi;
  alive[i] = 1;
  decided[i] = 0;
}

__global__ void k_iota_i32(int n, int* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

__global__ void k_gather_flag(int m, const int* idx, const unsigned char* owned, int* flag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m) flag[t] = owned[idx[t]] ? 1 : 0;
}

__global__ void k_cos_row(int n, const double* dec, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = cos(__dmul_rn(dec[i], kDeg2Rad));
}

// ---- tracker kill targets (fof.py:238-246, Q3: column 4, first occurrence) ------------------------------
__global__ void k_col4_keys(View Cn, int m, double* key, int* val) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  key[i] = Cn.at(i, 4) + 0.0;  // -0.0 -> +0.0 so equal values sort together
  val[i] = i;
}

__global__ void k_kill_targets(View G, int n, int m, const double* skey, const int* sval, int* target, int* t_count,
                               int* dup) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const double v = G.at(g, 4) + 0.0;
  int lo = 0, hi = m;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (skey[mid] < v) lo = mid + 1; else hi = mid;
  }
  const int tgt = (lo < m && skey[lo] == v) ? sval[lo] : -1;  // stable sort => smallest centre index first
  target[g] = tgt;
  if (tgt >= 0 && atomicAdd(t_count + tgt, 1) > 0) atomicAdd(dup, 1);
}

// Every galaxy row that carries centre i's column-4 value switches i off when it joins a cluster (np.intersect1d
// matches by VALUE, fof.py:238-246): t_rows[t_off[i] .. t_off[i+1]) lists those rows.  With unique column-4 values the
// list has one entry; photo-z catalogues quote z_upper to a few decimals and have hundreds.
__global__ void k_kill_rows(int n, const int* target, const long long* t_off, int* cursor, int* t_rows) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int tgt = target[g];
  if (tgt >= 0) t_rows[t_off[tgt] + atomicAdd(cursor + tgt, 1)] = g;
}

__global__ void k_i32_to_ll(int n, const int* in, long long* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// ---- centre cell list for the blocking test ----------------------------------------------------------------
__global__ void k_cc_keys(int m, const int* order, const double* ra, const double* dec, double ra0, double dec0,
                          double inv_sra, double inv_sdec, int ncx, int ncy, unsigned* key) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const int i = order[t];
  int cx = (int)floor((ra[i] - ra0) * inv_sra), cy = (int)floor((dec[i] - dec0) * inv_sdec);
  cx = min(max(cx, 0), ncx - 1);
  cy = min(max(cy, 0), ncy - 1);
  key[t] = (unsigned)(cy * ncx + cx);
}

__global__ void k_gather_d(int m, const int* idx, const double* src, double* dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < m) dst[t] = src[idx[t]];
}

__global__ void k_cc_starts(int m, const unsigned* key, int ncell, int* start) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > m) return;
  const int cur = k < m ? (int)key[k] : ncell;
  const int prev = k > 0 ? (int)key[k - 1] : -1;
  for (int c = prev + 1; c <= cur; ++c) start[c] = k;
}

struct CentreCells {
  double ra0, dec0, inv_sra, inv_sdec;
  int ncx, ncy;
  const int* start;
  const int* idx;     // centre index, cells sorted by centre redshift
  const double* z;    // centre redshift in the same order
  // the same order, packed for the blocking scan: a = (ra, dec, z, centre index as bits), b = (cos dec, R1, T1, -).
  // A scan walks consecutive entries, so the candidates' scalars arrive by whole cache lines instead of one random
  // sector per per-centre array.
  const double4* a;
  const double4* b;
};

__global__ void k_cc_pack(int m, const int* idx, const double* c_ra, const double* c_dec, const double* c_cos, const double* c_z,
                          const double* c_R1, const double* c_T1, double4* a, double4* b) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const int j = idx[k];
  a[k] = make_double4(c_ra[j], c_dec[j], c_z[j], __longlong_as_double((long long)j));
  b[k] = make_double4(c_cos[j], c_R1[j], c_T1[j], 0.0);
}

// ---- K9a: which active centres can be evaluated now ----------------------------------------------------------
// Centre j (higher priority: j < i) can claim the galaxy that switches centre i off only if that galaxy lies
// in j's velocity window and within j's virial search radius (members are a subset of virial_points,
// fof.py:180-228).  i is ready when no alive, undecided such j exists.  Cells are z-sorted, so only the
// centres whose window can contain the galaxy are visited.
// The scan taken cell by cell and candidate by candidate: the least work when a blocker turns up early, as it does
// for almost every centre of the first rounds (two million threads hide each other's latency).
__global__ void __launch_bounds__(128, 8) k_ready_lazy(int na, const int* active, CentreCells cc, const double* c_ra, const double* c_dec,
                        const double* c_cos, const double* c_z, const double* c_R1, const double* c_T1,
                        const unsigned char* alive, const unsigned char* decided, const long long* t_off, const int* t_rows,
                        const double* ra_row, const double* dec_row,
                        const double* cos_row, const double* z_row, double vmax, double R1max, int* ready_flag,
                        int* resume_row, int* resume_cell, int* resume_k) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= na) return;
  const int i = active[t];
  // centre i is switched off when a galaxy row carrying its column-4 value joins a cluster (Q3)
  const int nrows = (int)(t_off[i + 1] - t_off[i]);
  if (nrows == 0) { ready_flag[t] = 1; return; }
  const int* rows = t_rows + t_off[i];
  // |v(z; z_j)| <= vmax  =>  (z - beta)/(1 + beta) <= z_j <= (z + beta)/(1 - beta), beta = vmax/c (+ margin)
  const double beta = vmax / kCkms;
  // A candidate that does not block now never will again (alive only falls, decided only rises), so the scan resumes
  // where the previous round found its blocker instead of starting over: one pass per centre over all rounds.  The
  // rows are walked in list order; a row whose scan ended without a blocker is done for good.
  int c0 = resume_cell[i], k0 = resume_k[i];
  bool blocked = false;
  for (int r = resume_row[i]; r < nrows && !blocked; ++r) {
    const int g = rows[r];
    const double ra = ra_row[g], dec = dec_row[g], cs = cos_row[g], z = z_row[g];
    const double zlo = (z - beta) / (1.0 + beta) - 1e-9, zhi = (z + beta) / (1.0 - beta) + 1e-9;
    const double reach = ra_reach(R1max, dec);
    int x0 = (int)floor((ra - reach - cc.ra0) * cc.inv_sra), x1 = (int)floor((ra + reach - cc.ra0) * cc.inv_sra);
    int y0 = (int)floor((dec - R1max * 1.000001 - cc.dec0) * cc.inv_sdec),
        y1 = (int)floor((dec + R1max * 1.000001 - cc.dec0) * cc.inv_sdec);
    x0 = max(x0, 0); y0 = max(y0, 0);
    x1 = min(x1, cc.ncx - 1); y1 = min(y1, cc.ncy - 1);
    const int nx = x1 - x0 + 1, ncells = nx * (y1 - y0 + 1);
    for (int cidx = c0 > 0 ? c0 : 0; cidx < ncells && !blocked; ++cidx) {
      const int yy = y0 + cidx / nx, xx = x0 + cidx % nx;
      int lo = cc.start[yy * cc.ncx + xx];
      const int e = cc.start[yy * cc.ncx + xx + 1];
      if (cidx == c0) {
        lo = k0;
      } else {
        int hi = e;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (cc.z[mid] < zlo) lo = mid + 1; else hi = mid; }
      }
      for (int k = lo; k < e; ++k) {
        const double4 ca = cc.a[k];
        if (ca.z > zhi) break;
        const int j = (int)__double_as_longlong(ca.w);
        if (j >= i || !alive[j] || decided[j]) continue;
        if (fabs(redshift_to_velocity(z, ca.z)) > vmax) continue;
        const double4 cb = cc.b[k];
        BubbleTest bt;
        bt.init_T(ca.x, ca.y, cb.x, cb.y, cb.z);
        if (bt.inside(ra, dec, cs)) {
          blocked = true;
          resume_row[i] = r;
          resume_cell[i] = cidx;
          resume_k[i] = k;
          break;
        }
      }
    }
    c0 = -1;  // the stored scan position belongs to the row it was found on
  }
  ready_flag[t] = blocked ? 0 : 1;
}


// The scan is a chain of dependent loads (cell table -> binary search -> candidate record -> its tracker bytes), and a
// thread runs it alone: what bounds the kernel is memory latency, not bandwidth or issue slots.  So the loads are
// issued in independent groups: the binary searches of all (up to 9) neighbour cells advance in lockstep, and the
// candidates of a cell are fetched four at a time, records first, then their tracker bytes, before any is judged.
// one copy of the bubble test for the readiness scan (kept out of line on purpose)
__device__ __noinline__ bool bubble_inside(double cra, double cdec, double ccos, double R1, double T1, double ra, double dec, double cs) {
  BubbleTest bt;
  bt.init_T(cra, cdec, ccos, R1, T1);
  return bt.inside(ra, dec, cs);
}

constexpr int kReadyCells = 3, kReadyBatch = 4, kReadyThreads = 128;

__global__ void k_ready(int na, const int* active, CentreCells cc, const double* c_ra, const double* c_dec,
                        const double* c_cos, const double* c_z, const double* c_R1, const double* c_T1,
                        const unsigned char* alive, const unsigned char* decided, const long long* t_off, const int* t_rows,
                        const double* ra_row, const double* dec_row,
                        const double* cos_row, const double* z_row, double vmax, double R1max, int* ready_flag,
                        int* resume_row, int* resume_cell, int* resume_k) {
  __shared__ int s_lo[kReadyCells][kReadyThreads], s_e[kReadyCells][kReadyThreads];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= na) return;
  const int i = active[t];
  // centre i is switched off when a galaxy row carrying its column-4 value joins a cluster (Q3)
  const int nrows = (int)(t_off[i + 1] - t_off[i]);
  if (nrows == 0) { ready_flag[t] = 1; return; }
  const int* rows = t_rows + t_off[i];
  // |v(z; z_j)| <= vmax  =>  (z - beta)/(1 + beta) <= z_j <= (z + beta)/(1 - beta), beta = vmax/c (+ margin)
  const double beta = vmax / kCkms;
  // A candidate that does not block now never will again (alive only falls, decided only rises), so the scan resumes
  // where the previous round found its blocker instead of starting over: one pass per centre over all rounds.  The
  // rows are walked in list order; a row whose scan ended without a blocker is done for good.
  int c0 = resume_cell[i], k0 = resume_k[i];
  bool blocked = false;
  for (int r = resume_row[i]; r < nrows && !blocked; ++r) {
    const int g = rows[r];
    const double ra = ra_row[g], dec = dec_row[g], cs = cos_row[g], z = z_row[g];
    const double zlo = (z - beta) / (1.0 + beta) - 1e-9, zhi = (z + beta) / (1.0 - beta) + 1e-9;
    const double reach = ra_reach(R1max, dec);
    int x0 = (int)floor((ra - reach - cc.ra0) * cc.inv_sra), x1 = (int)floor((ra + reach - cc.ra0) * cc.inv_sra);
    int y0 = (int)floor((dec - R1max * 1.000001 - cc.dec0) * cc.inv_sdec),
        y1 = (int)floor((dec + R1max * 1.000001 - cc.dec0) * cc.inv_sdec);
    x0 = max(x0, 0); y0 = max(y0, 0);
    x1 = min(x1, cc.ncx - 1); y1 = min(y1, cc.ncy - 1);
    const int nx = x1 - x0 + 1, ncells = nx * (y1 - y0 + 1);
    // cells are taken in groups of kReadyCells; a stored scan position is the first cell of its group
    for (int cbase = c0 > 0 ? c0 : 0; cbase < ncells && !blocked; cbase += kReadyCells) {
      // candidate range [lo, e) of each cell of the group: entries with zlo <= z_j (the upper end is checked on the way)
      int lo[kReadyCells], e[kReadyCells];
#pragma unroll
      for (int u = 0; u < kReadyCells; ++u) {
        const int cidx = cbase + u;
        lo[u] = e[u] = 0;
        if (cidx < ncells) {
          const int cell = (y0 + cidx / nx) * cc.ncx + x0 + cidx % nx;
          lo[u] = cc.start[cell];
          e[u] = cc.start[cell + 1];
        }
      }
      int hi[kReadyCells];
#pragma unroll
      for (int u = 0; u < kReadyCells; ++u) {
        hi[u] = e[u];
        if (cbase + u == c0) { lo[u] = k0; hi[u] = k0; }  // the stored scan position: no search
      }
      for (bool any = true; any;) {
        any = false;
#pragma unroll
        for (int u = 0; u < kReadyCells; ++u)
          if (lo[u] < hi[u]) {
            const int mid = (lo[u] + hi[u]) >> 1;
            if (cc.z[mid] < zlo) lo[u] = mid + 1; else hi[u] = mid;
            any = true;
          }
      }
      // the ranges go through shared memory so that the candidate loop below exists once (not once per cell: the bubble
      // test is a large body, and copies of it thrash the instruction cache)
#pragma unroll
      for (int u = 0; u < kReadyCells; ++u) {
        s_lo[u][threadIdx.x] = lo[u];
        s_e[u][threadIdx.x] = e[u];
      }
#pragma unroll 1
      for (int u = 0; u < kReadyCells && !blocked; ++u) {
        const int ce = s_e[u][threadIdx.x];
#pragma unroll 1
        for (int k = s_lo[u][threadIdx.x]; k < ce && !blocked; k += kReadyBatch) {
          double4 ca[kReadyBatch];
          bool live[kReadyBatch];
#pragma unroll
          for (int v = 0; v < kReadyBatch; ++v) ca[v] = k + v < ce ? cc.a[k + v] : make_double4(0.0, 0.0, INFINITY, 0.0);
#pragma unroll
          for (int v = 0; v < kReadyBatch; ++v) {
            const int j = (int)__double_as_longlong(ca[v].w);
            live[v] = ca[v].z <= zhi && j < i && alive[j] && !decided[j];
          }
          // which of the four is the first that blocks (or ends the cell)?  The bubble test itself is not unrolled.
          bool past = false;
          int hit = -1;
#pragma unroll
          for (int v = 0; v < kReadyBatch; ++v) {
            if (hit >= 0 || past) continue;
            if (ca[v].z > zhi) { past = true; continue; }  // cells are z-sorted: nothing further in this cell
            if (!live[v]) continue;
            if (fabs(redshift_to_velocity(z, ca[v].z)) > vmax) continue;
            const double4 cb = cc.b[k + v];
            const double dd = fabs(ca[v].y - dec);
            if (dd > cb.y * 1.000001) continue;  // farther in declination than j's radius: cannot be inside
            if (bubble_inside(ca[v].x, ca[v].y, cb.x, cb.y, cb.z, ra, dec, cs)) hit = v;
          }
          if (hit >= 0) {
            blocked = true;
            resume_row[i] = r;
            resume_cell[i] = cbase + u;
            resume_k[i] = k + hit;
          }
          if (past) break;
        }
      }
    }
    c0 = -1;  // the stored scan position belongs to the row it was found on
  }
  ready_flag[t] = blocked ? 0 : 1;
}

// Same test with one warp per centre and one lane per neighbour cell: used for the late, small rounds, where the
// serial chain of dependent loads of a single thread is what the round waits for.
__global__ void k_ready_warp(int na, const int* active, CentreCells cc, const double* c_ra, const double* c_dec,
                             const double* c_cos, const double* c_z, const double* c_R1, const double* c_T1,
                             const unsigned char* alive, const unsigned char* decided, const long long* t_off,
                             const int* t_rows, const double* ra_row, const double* dec_row, const double* cos_row,
                             const double* z_row, double vmax, double R1max, int* ready_flag) {
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t >= na) return;
  const int i = active[t];
  const int nrows = (int)(t_off[i + 1] - t_off[i]);
  const int* rows = t_rows + t_off[i];
  const double beta = vmax / kCkms;
  bool blocked = false;
  for (int r = 0; r < nrows && !blocked; ++r) {
    const int g = rows[r];
    const double ra = ra_row[g], dec = dec_row[g], cs = cos_row[g], z = z_row[g];
    const double zlo = (z - beta) / (1.0 + beta) - 1e-9, zhi = (z + beta) / (1.0 - beta) + 1e-9;
    const double reach = ra_reach(R1max, dec);
    int x0 = (int)floor((ra - reach - cc.ra0) * cc.inv_sra), x1 = (int)floor((ra + reach - cc.ra0) * cc.inv_sra);
    int y0 = (int)floor((dec - R1max * 1.000001 - cc.dec0) * cc.inv_sdec),
        y1 = (int)floor((dec + R1max * 1.000001 - cc.dec0) * cc.inv_sdec);
    x0 = max(x0, 0); y0 = max(y0, 0);
    x1 = min(x1, cc.ncx - 1); y1 = min(y1, cc.ncy - 1);
    const int nx = x1 - x0 + 1, ncells = nx * (y1 - y0 + 1);
    for (int cidx = lane; cidx < ncells && !blocked; cidx += 32) {
      const int yy = y0 + cidx / nx, xx = x0 + cidx % nx;
      int lo = cc.start[yy * cc.ncx + xx];
      const int e = cc.start[yy * cc.ncx + xx + 1];
      int hi = e;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (cc.z[mid] < zlo) lo = mid + 1; else hi = mid; }
      for (int k = lo; k < e; ++k) {
        const double4 ca = cc.a[k];
        if (ca.z > zhi) break;
        const int j = (int)__double_as_longlong(ca.w);
        if (j >= i || !alive[j] || decided[j]) continue;
        if (fabs(redshift_to_velocity(z, ca.z)) > vmax) continue;
        const double4 cb = cc.b[k];
        BubbleTest bt;
        bt.init_T(ca.x, ca.y, cb.x, cb.y, cb.z);
        if (bt.inside(ra, dec, cs)) { blocked = true; break; }
      }
    }
    blocked = __any_sync(0xffffffffu, blocked);
  }
  if (lane == 0) ready_flag[t] = blocked ? 0 : 1;
}

__global__ void k_still_active(int na, const int* active, const unsigned char* alive, const unsigned char* decided,
                               int* flag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= na) return;
  const int i = active[t];
  flag[t] = (alive[i] && !decided[i]) ? 1 : 0;
}

// ---- K7: virial bubble of a ready centre, one warp per centre ---------------------------------------------------
// mode 0: count only (also records the window bbox); mode 1: emit keys (G1 cell key << 32 | row).
template <int MODE>
__global__ void k_virial(int nr, const int* ready, CatalogView cat, CellListView cl, const double* c_ra,
                         const double* c_dec, const double* c_cos, const double* c_R1, const int* c_wlo,
                         const int* c_whi, int grid_cells, int min_virial, int* nv, double4* bbox1,
                         const long long* voff, int* cursor, unsigned long long* keys) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nr) return;
  const int c = ready[w];
  const int wlo = c_wlo[c], whi = c_whi[c];
  if (MODE == 1 && nv[w] < min_virial) return;
  int count = 0;
  if (whi > wlo) {
    const double ra = c_ra[c], dec = c_dec[c], R1 = c_R1[c];
    BBox bb;
    Box box;
    int oof = 0;
    if (lane == 0) {
      bb = range_bbox(cat, wlo, whi);
      box = grid_box(bb, grid_cells, ra, dec, R1);
      // a single centre farther than r outside the grid's bounding box has no neighbours (App. B.1)
      oof = (ra - R1 > bb.ra_max + 1e-6) || (ra + R1 < bb.ra_min - 1e-6) || (dec - R1 > bb.dec_max + 1e-6) ||
            (dec + R1 < bb.dec_min - 1e-6);
      if (MODE == 0) bbox1[w] = make_double4(bb.ra_min, bb.ra_max, bb.dec_min, bb.dec_max);
    }
    bb.ra_min = __shfl_sync(0xffffffffu, bb.ra_min, 0); bb.ra_max = __shfl_sync(0xffffffffu, bb.ra_max, 0);
    bb.dec_min = __shfl_sync(0xffffffffu, bb.dec_min, 0); bb.dec_max = __shfl_sync(0xffffffffu, bb.dec_max, 0);
    box.ra_lo = __shfl_sync(0xffffffffu, box.ra_lo, 0); box.ra_hi = __shfl_sync(0xffffffffu, box.ra_hi, 0);
    box.dec_lo = __shfl_sync(0xffffffffu, box.dec_lo, 0); box.dec_hi = __shfl_sync(0xffffffffu, box.dec_hi, 0);
    oof = __shfl_sync(0xffffffffu, oof, 0);
    if (!oof) {
      BubbleTest bt;
      bt.init(ra, dec, c_cos[c], R1);
      const double reach = ra_reach(R1, dec);
      const int x0 = cl.cx_of(ra - reach), x1 = cl.cx_of(ra + reach);
      const int y0 = cl.cy_of(dec - R1 * 1.000001), y1 = cl.cy_of(dec + R1 * 1.000001);
      const int nx = x1 - x0 + 1, ncells = nx * (y1 - y0 + 1);
      GridAxis gx, gy;
      gx.init(bb.ra_min, bb.ra_max, grid_cells);
      gy.init(bb.dec_min, bb.dec_max, grid_cells);
      // the lanes first find the in-window run of one neighbour cell each (the cells are few), then the whole warp
      // strides over the candidates of every run: a loop per lane over its own cell would keep 3 of 32 lanes busy
      for (int base = 0; base < ncells; base += 32) {
        const int t = base + lane;
        int first = 0, last = 0;
        if (t < ncells) {
          const int yy = y0 + t / nx, xx = x0 + t % nx;
          const int s = cl.cell_start[yy * cl.ncx + xx], e = cl.cell_start[yy * cl.ncx + xx + 1];
          int a = s, b = e;
          while (a < b) { const int mid = (a + b) >> 1; if (cl.zrank[mid] < wlo) a = mid + 1; else b = mid; }
          first = a;
          b = e;
          while (a < b) { const int mid = (a + b) >> 1; if (cl.zrank[mid] < whi) a = mid + 1; else b = mid; }
          last = a;
        }
        const int here = ncells - base < 32 ? ncells - base : 32;
        for (int k = 0; k < here; ++k) {
          const int f = __shfl_sync(0xffffffffu, first, k), e = __shfl_sync(0xffffffffu, last, k);
          for (int j0 = f; j0 < e; j0 += 32) {
            const int j = j0 + lane;
            bool hit = false;
            double pra = 0.0, pdec = 0.0;
            if (j < e) {
              pra = cl.ra[j];
              pdec = cl.dec[j];
              hit = box.contains(pra, pdec) && bt.inside(pra, pdec, cl.cosdec[j]);
            }
            if (MODE == 0) {
              count += hit ? 1 : 0;
            } else {
              const unsigned mh = __ballot_sync(0xffffffffu, hit);
              if (mh) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(cursor + w, __popc(mh));
                pos = __shfl_sync(0xffffffffu, pos, 0);
                if (hit)
                  keys[voff[w] + pos + __popc(mh & ((1u << lane) - 1u))] =
                      ((unsigned long long)(unsigned)(gy.digit_clamped(pdec) * grid_cells + gx.digit_clamped(pra)) << 32) |
                      (unsigned)cl.row[j];
              }
            }
          }
        }
      }
    }
  } else if (MODE == 0 && lane == 0) {
    bbox1[w] = make_double4(0, 0, 0, 0);
  }
  if (MODE == 0) {
    for (int o = 16; o; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if (lane == 0) nv[w] = count;
  }
}

__global__ void k_alloc_sizes(int nr, const int* nv, int min_virial, long long* need) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < nr) need[w] = nv[w] >= min_virial ? nv[w] : 0;
}

// ---- K8a: linking length, friends grid and hop 1 (fof.py:187-206), one warp per centre ------------------------
__global__ void k_hop1(int nr, const int* ready, const int* nv, const long long* voff, const unsigned long long* keys1,
                       const double* ra_row, const double* dec_row, const double* cos_row, const double* c_ra,
                       const double* c_dec, const double* c_cos, const double* c_z, Cosmo cosmo, fofb_params p,
                       double* LLout, double4* bbox2, int* nF, unsigned long long* keys2, int key_bits, int* seg_of) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nr) return;
  const int n_v = nv[w];
  if (n_v < p.min_virial) {
    if (lane == 0) nF[w] = 0;
    return;
  }
  const int c = ready[w];
  const long long off = voff[w];
  const double zc = c_z[c], ra = c_ra[c], dec = c_dec[c];
  const double mean_sep = cosmo_mean_separation(cosmo, (double)n_v, zc, p.max_velocity, p.survey_area);
  const double LL = cosmo_proper_mpc_to_deg(cosmo, p.linking_length_factor * mean_sep, zc);
  // bbox of virial_points = the grid f_gsp is built on (fof.py:202)
  double4 bb = make_double4(INFINITY, -INFINITY, INFINITY, -INFINITY);
  for (int k = lane; k < n_v; k += 32) {
    const int row = (int)(unsigned)(keys1[off + k] & 0xffffffffull);
    const double pra = ra_row[row], pdec = dec_row[row];
    bb = bbox_merge(bb, make_double4(pra, pra, pdec, pdec));
  }
  for (int o = 16; o; o >>= 1) {
    double4 other;
    other.x = __shfl_xor_sync(0xffffffffu, bb.x, o); other.y = __shfl_xor_sync(0xffffffffu, bb.y, o);
    other.z = __shfl_xor_sync(0xffffffffu, bb.z, o); other.w = __shfl_xor_sync(0xffffffffu, bb.w, o);
    bb = bbox_merge(bb, other);
  }
  const BBox b2{bb.x, bb.y, bb.z, bb.w};
  const bool oof = (ra - LL > b2.ra_max + 1e-6) || (ra + LL < b2.ra_min - 1e-6) || (dec - LL > b2.dec_max + 1e-6) ||
                   (dec + LL < b2.dec_min - 1e-6);
  const Box box = grid_box(b2, p.grid_cells, ra, dec, LL);
  BubbleTest bt;
  bt.init(ra, dec, c_cos[c], LL);
  GridAxis gx, gy;
  gx.init(b2.ra_min, b2.ra_max, p.grid_cells);
  gy.init(b2.dec_min, b2.dec_max, p.grid_cells);
  int cnt = 0;
  for (int k = lane; k < n_v; k += 32) {
    const int row = (int)(unsigned)(keys1[off + k] & 0xffffffffull);
    const double pra = ra_row[row], pdec = dec_row[row];
    const bool fr = !oof && box.contains(pra, pdec) && bt.inside(pra, pdec, cos_row[row]);
    cnt += fr ? 1 : 0;
    const unsigned long long cellkey = (unsigned)(gy.digit_clamped(pdec) * p.grid_cells + gx.digit_clamped(pra));
    // key_bits == 0: keys1 is in virial order and k is the tie-break inside a cell of the friends' grid.
    // key_bits  > 0: keys1 is unsorted; its own (cell of the virial grid, row) is that tie-break, so ONE sort by
    //                (friend flag, friends' cell, virial cell, row) gives the same order without sorting keys1 first.
    keys2[off + k] = key_bits ? (fr ? 0ull : (1ull << 63)) | (cellkey << (32 + key_bits)) | keys1[off + k]
                              : (fr ? 0ull : (1ull << 63)) | (cellkey << 32) | (unsigned)k;
    if (seg_of) seg_of[off + k] = w;
  }
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) {
    nF[w] = cnt;
    LLout[w] = LL;
    bbox2[w] = bb;
  }
}

__global__ void k_hash_sizes(int nr, const int* nv, const int* nF, int min_virial, long long* cap) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= nr) return;
  long long c = 0;
  if (nv[w] >= min_virial && nF[w] > 0 && nF[w] < nv[w]) {
    c = 64;
    while (c < 16ll * nv[w]) c <<= 1;
  }
  cap[w] = c;
}

// ---- value set for the element-wise novelty rule (fof.py:217-223, Q5) --------------------------------------------
constexpr unsigned long long kHashEmpty = ~0ull;

__device__ __forceinline__ unsigned long long hash_mix(unsigned long long x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

__device__ __forceinline__ bool value_bits(double v, unsigned long long* bits) {
  if (v != v) return false;  // NaN equals nothing (np.isin)
  *bits = (unsigned long long)__double_as_longlong(v + 0.0);
  return true;
}

__device__ __forceinline__ void hash_insert(unsigned long long* tab, long long cap, double v) {
  unsigned long long b;
  if (!value_bits(v, &b)) return;
  long long s = (long long)(hash_mix(b) & (unsigned long long)(cap - 1));
  for (;;) {
    const unsigned long long old = atomicCAS(tab + s, kHashEmpty, b);
    if (old == kHashEmpty || old == b) return;
    s = (s + 1) & (cap - 1);
  }
}

__device__ __forceinline__ bool hash_contains(const unsigned long long* tab, long long cap, double v) {
  unsigned long long b;
  if (!value_bits(v, &b)) return false;
  long long s = (long long)(hash_mix(b) & (unsigned long long)(cap - 1));
  for (;;) {
    const unsigned long long cur = __ldcg(tab + s);
    if (cur == b) return true;
    if (cur == kHashEmpty) return false;
    s = (s + 1) & (cap - 1);
  }
}

// Single-owner variants for phase (d): only one CTA touches the table, so ordinary cached loads/stores with
// block barriers in between are coherent (same SM, same L1) and avoid an L2 round trip per probe.
__device__ __forceinline__ void hash_insert_owner(unsigned long long* tab, long long cap, double v) {
  unsigned long long b;
  if (!value_bits(v, &b)) return;
  long long s = (long long)(hash_mix(b) & (unsigned long long)(cap - 1));
  for (;;) {
    const unsigned long long cur = tab[s];
    if (cur == b) return;
    if (cur == kHashEmpty) { tab[s] = b; return; }
    s = (s + 1) & (cap - 1);
  }
}

__device__ __forceinline__ bool hash_contains_owner(const unsigned long long* tab, long long cap, double v) {
  unsigned long long b;
  if (!value_bits(v, &b)) return false;
  long long s = (long long)(hash_mix(b) & (unsigned long long)(cap - 1));
  for (;;) {
    const unsigned long long cur = tab[s];
    if (cur == b) return true;
    if (cur == kHashEmpty) return false;
    s = (s + 1) & (cap - 1);
  }
}

// ---- K8b: hop 2 and the richness gate (fof.py:208-246) -------------------------------------------------------
// A row is admitted in the batch of the FIRST friend whose bubble contains it, iff none of its values occurs in
// the member matrix as it stands before that batch; the value set only grows, so a row that fails once can never
// be admitted later, and a row whose values already collide with the hop-1 members is out from the start.
// Four launches: (a) V2 order and friend rows, warp/centre; (b) hop-1 member values into the per-centre hash set,
// thread/V2 entry; (c) novelty of every non-member against that set, thread/V2 entry; (d) the few surviving
// candidates are decided batch by batch, warp/centre.
__global__ void k_hop2_setup(int nr, const int* nv, const int* nF, const long long* voff, const unsigned long long* keys1,
                             const unsigned long long* keys2, const double* ra_row, const double* dec_row,
                             const double* cos_row, int min_virial, int* v2row, int* owner, double* v2ra, double* v2dec,
                             double* v2cos, int* mrows, int key_bits) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nr) return;
  const int n_v = nv[w];
  if (n_v < min_virial) return;
  const long long off = voff[w];
  const int n_f = nF[w];
  // V2 order: friends first (in f_gsp's return order), then the rest in f_gsp cell order
  for (int t = lane; t < n_v; t += 32) {
    const int low = (int)(unsigned)(keys2[off + t] & 0xffffffffull);  // the row itself, or the position in virial order
    const int row = key_bits ? low : (int)(unsigned)(keys1[off + low] & 0xffffffffull);
    v2row[off + t] = row;
    owner[off + t] = w;
    v2ra[off + t] = ra_row[row];
    v2dec[off + t] = dec_row[row];
    v2cos[off + t] = cos_row[row];
    if (t < n_f) mrows[off + t] = row;
  }
}

__global__ void k_hop2_insert(long long total, const int* owner, const long long* voff, const int* nF, const long long* hoff,
                              const long long* hcap, unsigned long long* htab, View G, const int* v2row) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const int w = owner[g];
  const long long cap = hcap[w];
  if (cap == 0 || g - voff[w] >= nF[w]) return;
  unsigned long long* tab = htab + hoff[w];
  const int row = v2row[g];
  for (int col = 0; col < G.cols; ++col) hash_insert(tab, cap, G.at(row, col));
}

__global__ void k_hop2_novel(long long total, const int* owner, const long long* voff, const int* nF, const long long* hoff,
                             const long long* hcap, const unsigned long long* htab, View G, const int* v2row,
                             unsigned char* novel) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total) return;
  const int w = owner[g];
  const long long cap = hcap[w];
  unsigned char nov = 0;
  if (cap > 0 && g - voff[w] >= nF[w]) {
    const unsigned long long* tab = htab + hoff[w];
    const int row = v2row[g];
    bool ok = true;  // columns from the last (N, then ID: the ones that collide in practice) to the first
    for (int col = G.cols - 1; col >= 0 && ok; --col) ok = !hash_contains(tab, cap, G.at(row, col));
    nov = ok ? 1 : 0;
  }
  novel[g] = nov;
}

// FOFB_RADIX_ABOVE=<n> overrides the number of virial points above which a round orders them with two device-wide
// radix sorts instead of one segmented sort (the parity tests are also run with 0 to cover that path)
static long long radix_sort_above() {
  static const long long v = [] {
    const char* e = getenv("FOFB_RADIX_ABOVE");
    return e ? atoll(e) : (1ll << 17);
  }();
  return v;
}

constexpr int kHop2Threads = 256;
constexpr int kHop2BitonicAbove = 1024;  // candidate count above which hop 2 sorts instead of ranking by counting

// FOFB_HOP2_SORT_ABOVE=<n> overrides the threshold (the parity tests are also run with 0 to cover the sorting path)
static int hop2_sort_above() {
  static const int v = [] {
    const char* e = getenv("FOFB_HOP2_SORT_ABOVE");
    return e ? atoi(e) : kHop2BitonicAbove;
  }();
  return v;
}

// FOFB_HOP2_SMEM_ROWS=<n> caps the number of hop-2 candidates the shared-memory path takes (the parity tests also run
// with 0, which sends every centre through the global-memory path)
static int hop2_smem_rows() {
  static const int v = [] {
    const char* e = getenv("FOFB_HOP2_SMEM_ROWS");
    return e ? atoi(e) : (1 << 30);
  }();
  return v;
}

// The richness gate and the tracker for one centre (fof.py:230-246): members that are candidate centres never seed a
// cluster later.  Called by all `nthreads` threads that share the centre (a warp or a CTA).
__device__ __forceinline__ void hop2_finalize(int w, int c, int n_members, long long off, int tid, int nthreads,
                                              const fofb_params& p, const int* mrows, const int* kill_target,
                                              unsigned char* alive, unsigned char* decided, int* nM, int* pass) {
  const bool ok = n_members >= p.richness;
  if (tid == 0) {
    nM[w] = n_members;
    pass[w] = ok ? 1 : 0;
    decided[c] = 1;
  }
  if (ok)
    for (int t = tid; t < n_members; t += nthreads) {
      const int tgt = kill_target[mrows[off + t]];
      if (tgt > c) alive[tgt] = 0;
    }
}

// (d1) one warp per centre: the novel non-members in V2 order (the hop-2 candidates).  Centres without candidates are
// finished here -- the common case: below ~100 virial points the linking length exceeds the virial radius and hop 1
// already holds every point.  The others are queued by size class for k_hop2_heavy.
__global__ void k_hop2_front(int nr, const int* ready, const int* nv, const int* nF, const long long* voff,
                             const long long* hcap, const unsigned char* novel, fofb_params p, int* nm_list, int* n_nm_out,
                             const int* mrows, int* nM, int* pass, const int* kill_target, unsigned char* alive,
                             unsigned char* decided, int small_rows, int large_rows, int* qcount, int* qlist) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nr) return;
  const int c = ready[w];
  const int n_v = nv[w];
  int n_members = 0, n_nm = 0;
  long long off = 0;
  if (n_v >= p.min_virial) {
    off = voff[w];
    const int n_f = nF[w];
    n_members = n_f;
    if (hcap[w] > 0) {
      for (int base = n_f; base < n_v; base += 32) {
        const int t = base + lane;
        const bool nov = t < n_v && novel[off + t];
        const unsigned mk = __ballot_sync(0xffffffffu, nov);
        if (nov) nm_list[off + n_nm + __popc(mk & ((1u << lane) - 1u))] = t;
        n_nm += __popc(mk);
      }
    }
  }
  if (lane == 0) n_nm_out[w] = n_nm;
  if (n_nm == 0) {
    hop2_finalize(w, c, n_members, off, lane, 32, p, mrows, kill_target, alive, decided, nM, pass);
  } else if (lane == 0) {
    const int cls = n_nm <= small_rows ? 0 : (n_nm <= large_rows ? 1 : 2);
    qlist[(size_t)cls * nr + atomicAdd(qcount + cls, 1)] = w;
  }
}

// Global-memory path of one centre for a 256-thread CTA (candidate sets that do not fit shared memory: the percolating
// blobs of the dense stress configuration).  Returns the member count.
__device__ int hop2_global_path(int w, int n_f, int n_nm, long long off, long long cap, unsigned long long* tab, View G,
                                const double* v2ra, const double* v2dec, const double* v2cos, double LL, double4 b4,
                                const fofb_params& p, const int* v2row, int* mrows, const int* nm_list, int* first_batch,
                                int* nm_sorted, unsigned long long* sort_scratch, int sort_above) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kHop2Threads / 32;
  __shared__ int s_members;
  int n_members = n_f;
  const BBox b2{b4.x, b4.y, b4.z, b4.w};
  const double sh = sin(0.5 * LL * kDeg2Rad);
  const double T = sh * sh;
  // first friend (in order) whose bubble contains each candidate.  Friends sit in the lanes, 32 per warp
  // step; the candidates stream past them; atomicMin merges the warps' findings.
  for (int q = tid; q < n_nm; q += kHop2Threads) first_batch[off + q] = 0x7fffffff;
  __syncthreads();
  for (int base = warp * 32; base < n_f; base += kHop2Threads) {
    const int fi = base + lane;
    const bool have = fi < n_f;
    const double fra = have ? v2ra[off + fi] : 0.0, fdec = have ? v2dec[off + fi] : 0.0;
    BubbleTest bt;
    bt.init_T(fra, fdec, have ? v2cos[off + fi] : 1.0, LL, T);
    const Box fbox = grid_box(b2, p.grid_cells, fra, fdec, LL);  // the friend's cell box: once per friend
    for (int q = 0; q < n_nm; ++q) {
      // an earlier friend already has it?  ONE lane reads, so that the whole warp takes the same branch: other warps
      // lower first_batch[q] concurrently, and lanes that re-converge late could otherwise read different values and
      // split over the ballot below
      int cur = 0;
      if (lane == 0) cur = *(volatile int*)(first_batch + off + q);
      cur = __shfl_sync(0xffffffffu, cur, 0);
      if (cur <= base) continue;
      const int vp = nm_list[off + q];
      const double pra = v2ra[off + vp], pdec = v2dec[off + vp];
      const bool hit = have && fbox.contains(pra, pdec) && bt.inside(pra, pdec, v2cos[off + vp]);
      const unsigned mh = __ballot_sync(0xffffffffu, hit);
      if (mh && lane == 0) atomicMin(first_batch + off + q, base + __ffs(mh) - 1);
    }
  }
  __syncthreads();
  // candidates ordered by (first batch, V2 position): rank by all-pairs counting -- or, for the tens of
  // thousands of candidates of a giant bubble, a bitonic sort of the unique keys (first batch, position) in
  // scratch memory (flip / disperse network: every exchange is ascending, so the virtual +inf padding up to
  // the next power of two never moves and out-of-range partners are simply skipped)
  if (n_nm > sort_above) {
    unsigned long long* key = sort_scratch + off;
    for (int q = tid; q < n_nm; q += kHop2Threads)
      key[q] = ((unsigned long long)(unsigned)first_batch[off + q] << 32) | (unsigned)q;
    __syncthreads();
    for (int k = 2; (k >> 1) < n_nm; k <<= 1) {
      for (int j = k; j > 1; j >>= 1) {
        const int mask = j == k ? k - 1 : (j >> 1);  // flip on the first step of a merge, disperse after
        for (int i = tid; i < n_nm; i += kHop2Threads) {
          const int l = i ^ mask;
          if (l > i && l < n_nm) {
            const unsigned long long a = key[i], b = key[l];
            if (a > b) { key[i] = b; key[l] = a; }
          }
        }
        __syncthreads();
      }
    }
    for (int q = tid; q < n_nm; q += kHop2Threads) nm_sorted[off + q] = (int)(unsigned)(key[q] & 0xffffffffull);
  } else
  for (int q = tid; q < n_nm; q += kHop2Threads) {
    const int fq = first_batch[off + q];
    int rank = 0;
    for (int r = 0; r < n_nm; ++r) {
      const int fr = first_batch[off + r];
      rank += (fr < fq || (fr == fq && r < q)) ? 1 : 0;
    }
    nm_sorted[off + rank] = q;
  }
  __syncthreads();
  // one pass over the batches in friend order; every candidate is decided in its first batch, against the
  // member matrix as it stood before that batch.
  const int cols = G.cols;
  int s0 = 0;
  while (s0 < n_nm) {
    const int fb = first_batch[off + nm_sorted[off + s0]];
    if (fb == 0x7fffffff) break;  // no friend's bubble contains the remaining rows
    int s1 = s0 + 1;
    while (s1 < n_nm && first_batch[off + nm_sorted[off + s1]] == fb) ++s1;
    __syncthreads();
    // pass 1: every row of the batch against the value set as it stands BEFORE the batch (a warp per row)
    for (int j = s0 + warp; j < s1; j += kWarps) {
      const int q = nm_sorted[off + j];
      const int row = v2row[off + nm_list[off + q]];
      bool found = false;
      for (int col = lane; col < cols && !found; col += 32) found = hash_contains_owner(tab, cap, G.at(row, col));
      const bool adm = !__any_sync(0xffffffffu, found);
      if (lane == 0 && adm) first_batch[off + q] = -1;
    }
    __syncthreads();
    // pass 2: append the admitted rows in order and extend the value set (warp 0)
    if (warp == 0) {
      int nm = n_members;
      for (int j = s0; j < s1; ++j) {
        const int q = nm_sorted[off + j];
        if (first_batch[off + q] != -1) continue;
        const int row = v2row[off + nm_list[off + q]];
        if (lane == 0) {  // one thread inserts: plain stores, no two writers
          mrows[off + nm] = row;
          for (int col = 0; col < cols; ++col) hash_insert_owner(tab, cap, G.at(row, col));
        }
        ++nm;
      }
      if (lane == 0) s_members = nm;
      __threadfence_block();
    }
    __syncthreads();
    n_members = s_members;
    s0 = s1;
  }
  __syncthreads();
  return n_members;
}

// (d2) the centres that have hop-2 candidates, one CTA at a time from the queue of a size class (persistent CTAs).
// Only collisions BETWEEN candidates are still open (every candidate is already novel against the hop-1 members), so
// the value set that matters is the candidates' own: their CAP x 8 values get a slot each in a shared-memory hash
// table (parallel insert), and the batch loop -- sequential by nature: a batch sees what earlier batches admitted --
// only reads and sets one byte per value.  Warp 0 walks the batches, four rows (x 8 columns) per step.
//   shared memory per candidate row: ra, dec, cos (24) + 16 table slots (128) + first batch, order, row (12)
//                                    + 8 slot ids (16) + 16 taken bytes = 196 bytes
constexpr int kHop2SmallRows = 256, kHop2LargeRows = 1024;
constexpr size_t hop2_smem_bytes(int cap_rows) { return (size_t)cap_rows * 196; }
constexpr unsigned long long kSlotEmpty = ~0ull;

template <int CAP>
__global__ void __launch_bounds__(kHop2Threads) k_hop2_heavy(
    int cls, int nr, const int* qcount, const int* qlist, const int* ready, const int* nF, const int* n_nm_in,
    const long long* voff, const long long* hoff, const long long* hcap, unsigned long long* htab, View G,
    const double* v2ra, const double* v2dec, const double* v2cos, const double* LLin, const double4* bbox2, fofb_params p,
    const int* v2row, int* mrows, const int* nm_list, int* first_batch, int* nm_sorted, int* nM, int* pass,
    const int* kill_target, unsigned char* alive, unsigned char* decided, unsigned long long* sort_scratch, int sort_above) {
  extern __shared__ __align__(16) unsigned char hop2_smem[];
  double* s_ra = reinterpret_cast<double*>(hop2_smem);
  double* s_dec = s_ra + CAP;
  double* s_cos = s_dec + CAP;
  unsigned long long* s_tab = reinterpret_cast<unsigned long long*>(s_cos + CAP);  // 16 * CAP slots
  int* s_fb = reinterpret_cast<int*>(s_tab + 16 * CAP);
  int* s_sorted = s_fb + CAP;
  int* s_row = s_sorted + CAP;
  unsigned short* s_slot = reinterpret_cast<unsigned short*>(s_row + CAP);  // 8 per row; 0xffff: NaN, matches nothing
  unsigned char* s_taken = reinterpret_cast<unsigned char*>(s_slot + 8 * CAP);
  __shared__ int s_nmem;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total = qcount[cls];
  for (int it = blockIdx.x; it < total; it += gridDim.x) {
    const int w = qlist[(size_t)cls * nr + it];
    const int c = ready[w];
    const long long off = voff[w];
    const int n_f = nF[w], n_nm = n_nm_in[w];
    const double LL = LLin[w];
    const double4 b4 = bbox2[w];
    int n_members;
    if (CAP == 0 || n_nm > CAP || G.cols != 8) {
      n_members = hop2_global_path(w, n_f, n_nm, off, hcap[w], htab + hoff[w], G, v2ra, v2dec, v2cos, LL, b4, p, v2row, mrows,
                                   nm_list, first_batch, nm_sorted, sort_scratch, sort_above);
    } else {
      const BBox b2{b4.x, b4.y, b4.z, b4.w};
      const double sh = sin(0.5 * LL * kDeg2Rad);
      const double T = sh * sh;
      int tcap = 64;
      while (tcap < 16 * n_nm) tcap <<= 1;
      __syncthreads();  // the previous centre of this CTA is done with the shared arrays
      for (int q = tid; q < n_nm; q += kHop2Threads) {
        const int vp = nm_list[off + q];
        s_ra[q] = v2ra[off + vp];
        s_dec[q] = v2dec[off + vp];
        s_cos[q] = v2cos[off + vp];
        s_row[q] = v2row[off + vp];
        s_fb[q] = 0x7fffffff;
      }
      for (int t = tid; t < tcap; t += kHop2Threads) {
        s_tab[t] = kSlotEmpty;
        s_taken[t] = 0;
      }
      __syncthreads();
      // every candidate value gets its slot (row q, column col <-> thread stride)
      for (int t = tid; t < 8 * n_nm; t += kHop2Threads) {
        unsigned long long bits;
        unsigned short slot = 0xffff;
        if (value_bits(G.at(s_row[t >> 3], t & 7), &bits)) {
          int sl = (int)(hash_mix(bits) & (unsigned long long)(tcap - 1));
          for (;;) {
            const unsigned long long old = atomicCAS(s_tab + sl, kSlotEmpty, bits);
            if (old == kSlotEmpty || old == bits) break;
            sl = (sl + 1) & (tcap - 1);
          }
          slot = (unsigned short)sl;
        }
        s_slot[t] = slot;
      }
      // first friend (in order) whose bubble contains each candidate: friends in the lanes, 32 per warp step, the
      // candidates stream past them out of shared memory
      for (int base = warp * 32; base < n_f; base += kHop2Threads) {
        const int fi = base + lane;
        const bool have = fi < n_f;
        const double fra = have ? v2ra[off + fi] : 0.0, fdec = have ? v2dec[off + fi] : 0.0;
        BubbleTest bt;
        bt.init_T(fra, fdec, have ? v2cos[off + fi] : 1.0, LL, T);
        const Box fbox = grid_box(b2, p.grid_cells, fra, fdec, LL);
        for (int q = 0; q < n_nm; ++q) {
          int cur = 0;  // one lane reads (see hop2_global_path): the skip must be warp-uniform
          if (lane == 0) cur = *(volatile int*)(s_fb + q);
          cur = __shfl_sync(0xffffffffu, cur, 0);
          if (cur <= base) continue;  // an earlier friend already has it
          const double pra = s_ra[q], pdec = s_dec[q];
          const bool hit = have && fbox.contains(pra, pdec) && bt.inside(pra, pdec, s_cos[q]);
          const unsigned mh = __ballot_sync(0xffffffffu, hit);
          if (mh && lane == 0) atomicMin(s_fb + q, base + __ffs(mh) - 1);
        }
      }
      __syncthreads();
      // order by (first batch, V2 position)
      for (int q = tid; q < n_nm; q += kHop2Threads) {
        const int fq = s_fb[q];
        int rank = 0;
        for (int r = 0; r < n_nm; ++r) {
          const int fr = s_fb[r];
          rank += (fr < fq || (fr == fq && r < q)) ? 1 : 0;
        }
        s_sorted[rank] = q;
      }
      __syncthreads();
      if (warp == 0) {
        int nm = n_f;
        int s0 = 0;
        const int sub = lane >> 3, col = lane & 7;
        while (s0 < n_nm) {
          const int fb = s_fb[s_sorted[s0]];
          if (fb == 0x7fffffff) break;  // no friend's bubble contains the remaining rows
          int s1 = s0 + 1;
          for (;;) {  // end of the batch
            const int j = s1 + lane;
            const bool differs = j >= n_nm || s_fb[s_sorted[j]] != fb;
            const unsigned mk = __ballot_sync(0xffffffffu, differs);
            if (mk) { s1 += __ffs(mk) - 1; break; }
            s1 += 32;
          }
          // every row of the batch against the admitted values as they stand BEFORE the batch
          for (int j0 = s0; j0 < s1; j0 += 4) {
            const int j = j0 + sub;
            bool blocked = false;
            int q = 0;
            if (j < s1) {
              q = s_sorted[j];
              const unsigned short sl = s_slot[8 * q + col];
              blocked = sl != 0xffff && s_taken[sl];
            }
            const unsigned mk = __ballot_sync(0xffffffffu, blocked);
            if (j < s1 && col == 0 && ((mk >> (lane & 24)) & 0xffu) == 0) s_fb[q] = -1;
          }
          __syncwarp();
          // append the admitted rows in order; their values join the set
          for (int j0 = s0; j0 < s1; j0 += 4) {
            const int j = j0 + sub;
            bool adm = false;
            int q = 0;
            if (j < s1) {
              q = s_sorted[j];
              adm = s_fb[q] == -1;
            }
            if (adm) {
              const unsigned short sl = s_slot[8 * q + col];
              if (sl != 0xffff) s_taken[sl] = 1;
            }
            const unsigned amk = __ballot_sync(0xffffffffu, adm && col == 0);
            if (adm && col == 0) mrows[off + nm + __popc(amk & ((1u << lane) - 1u))] = s_row[q];
            nm += __popc(amk);
          }
          __syncwarp();
          s0 = s1;
        }
        if (lane == 0) s_nmem = nm;
      }
      __syncthreads();
      n_members = s_nmem;
    }
    __threadfence_block();
    __syncthreads();  // mrows of this centre are complete before the tracker pass reads them
    hop2_finalize(w, c, n_members, off, tid, kHop2Threads, p, mrows, kill_target, alive, decided, nM, pass);
  }
}

__global__ void k_commit_sizes(int nr, const int* pass, const int* nM, long long* len) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < nr) len[w] = pass[w] ? nM[w] : 0;
}

__global__ void k_commit(int nr, const int* ready, const int* pass, const int* nM, const long long* voff,
                         const int* mrows, const long long* dst_off, const long long* slot_of, long long pool_base,
                         int pool_k0, int* pool_rows, int* pool_centre, long long* pool_off, int* pool_len) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nr || !pass[w]) return;
  const long long src = voff[w], dst = pool_base + dst_off[w];
  const int len = nM[w];
  for (int t = lane; t < len; t += 32) pool_rows[dst + t] = mrows[src + t];
  if (lane == 0) {
    const int slot = pool_k0 + (int)slot_of[w];
    pool_centre[slot] = ready[w];
    pool_off[slot] = dst;
    pool_len[slot] = len;
  }
}

__global__ void k_pass_to_ll(int nr, const int* pass, long long* out) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w < nr) out[w] = pass[w];
}

// candidates in priority order: gather from the pool following the sorted slot permutation
__global__ void k_list_sizes(int K, const int* slot, const int* pool_len, long long* len) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < K) len[k] = pool_len[slot[k]];
}

__global__ void k_list_gather(int K, const int* slot, const int* pool_centre, const long long* pool_off,
                              const int* pool_len, const int* pool_rows, const long long* off, int* centre, int* rows) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  const int s = slot[k];
  const long long src = pool_off[s], dst = off[k];
  for (int t = lane; t < pool_len[s]; t += 32) rows[dst + t] = pool_rows[src + t];
  if (lane == 0) centre[k] = pool_centre[s];
}

// exclusive scan of len[0..n) into off[0..n]; returns the total (host)
__global__ void k_scan_total_fs(const long long* len, long long* off, int n) { off[n] = off[n - 1] + len[n - 1]; }

static long long scan_offsets(fofb_handle* h, const long long* len, long long* off, int n) {
  if (n == 0) return 0;
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, len, off, n, h->stream));
  FOFB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp(h, bytes), bytes, len, off, n, h->stream));
  k_scan_total_fs<<<1, 1, 0, h->stream>>>(len, off, n);  // off[n] = total, on the device
  count_launch(h, 3);
  return d2h_scalar(h, off + n);
}

static int select_flagged(fofb_handle* h, const int* in, const int* flags, int* out, int n, int* d_count) {
  if (n == 0) return 0;
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, in, flags, out, d_count, n, h->stream));
  FOFB_CUDA(cub::DeviceSelect::Flagged(cub_tmp(h, bytes), bytes, in, flags, out, d_count, n, h->stream));
  count_launch(h, 2);
  return d2h_scalar(h, d_count);
}

static void segmented_sort(fofb_handle* h, const unsigned long long* in, unsigned long long* out, long long total,
                           int nseg, const long long* off) {
  if (total == 0 || nseg == 0) return;
  size_t bytes = 0;
  FOFB_CUDA(cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, in, out, total, nseg, off, off + 1, h->stream));
  FOFB_CUDA(cub::DeviceSegmentedSort::SortKeys(cub_tmp(h, bytes), bytes, in, out, total, nseg, off, off + 1, h->stream));
  count_launch(h, 4);
}

void fof_overlap_and_bcg(fofb_handle* h, FofState& S);  // overlap.cu

// wall-clock of the steps of one run, for FOFB_TRACE=1 (diagnostics only)
static double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

void FofState::run(fofb_handle* h, const fofb_params& p, const View& G_in, const View& Cn_in) {
  const bool trace = getenv("FOFB_TRACE") != nullptr;
  double t0 = trace ? now_ms() : 0.0;
  prepare(h, p, G_in, Cn_in, nullptr);
  if (trace) { cudaStreamSynchronize(h->stream); fprintf(stderr, "[trace] prepare %.2f ms (m=%d)\n", now_ms() - t0, m); t0 = now_ms(); }
  while (na > 0) {
    const int na0 = na;
    const long long ev0 = evaluated;
    round_evaluate(h);
    round_advance(h);
    if (trace) { cudaStreamSynchronize(h->stream); fprintf(stderr, "[trace] round %d: active %d evaluated %lld -> %.2f ms\n", rounds, na0, evaluated - ev0, now_ms() - t0); t0 = now_ms(); }
  }
  finalize(h);
  if (trace) { cudaStreamSynchronize(h->stream); fprintf(stderr, "[trace] finalize %.2f ms\n", now_ms() - t0); t0 = now_ms(); }
  fof_overlap_and_bcg(h, *this);
  if (trace) { cudaStreamSynchronize(h->stream); fprintf(stderr, "[trace] overlap+bcg %.2f ms (K=%d)\n", now_ms() - t0, candidates.K); }
  ran = true;
}

// Everything before the priority rounds.  `owned` (device, one byte per centre, may be null = all): the centres this
// process evaluates; the others take part only as potential blockers whose state arrives from their owner.
void FofState::prepare(fofb_handle* h, const fofb_params& p, const View& G_in, const View& Cn_in, const unsigned char* owned) {
  cudaStream_t st = h->stream;
  ran = finished = false;
  distributed = owned != nullptr;
  na = 0;
  params = p;
  G = G_in;
  Cn = Cn_in;
  FOFB_REQUIRE(G.cols >= 8 && Cn.cols >= 8, "FoF needs the 8-column layout (ra, dec, z, z_lower, z_upper, RMag, ID, N)");
  FOFB_REQUIRE(G.cols == Cn.cols, "galaxy_data and candidate_centers must have the same number of columns");
  FOFB_REQUIRE(G.cols <= 64, "too many columns");
  n = (int)G.rows;
  m = (int)Cn.rows;
  const int B = 256;
  const Cosmo cosmo = device_cosmo(h, p);
  Catalog& cat = cat_();
  if (catp == &cat_own) cat.build(h, G);
  FOFB_REQUIRE(cat.n == G.rows, "adopted catalogue does not match the FoF() rows");
  CatalogView cv = cat.view();
  k_cos_row<<<grid_for(n, B), B, 0, st>>>(n, cat.dec_row.p, cos_row.alloc(n));
  FOFB_LAUNCH_CHECK(h);
  candidates.K = merged.K = bcg.K = final_.K = 0;
  candidates.total = merged.total = bcg.total = final_.total = 0;
  pool_used = 0;
  pool_K = 0;
  rounds = 0;
  evaluated = 0;
  if (m == 0) return;
  const size_t mm = (size_t)m;
  FOFB_PROFILED(h, "k_centre_prep", k_centre_prep<<<grid_for(m, B), B, 0, st>>>(Cn, m, cv, cosmo, p.max_velocity, p.virial_radius, p.count_radius,
                                              c_ra.alloc(mm), c_dec.alloc(mm), c_cos.alloc(mm), c_z.alloc(mm),
                                              c_R1.alloc(mm), c_T1.alloc(mm), c_thvir.alloc(mm), c_th05.alloc(mm), c_mag.alloc(mm),
                                              c_N.alloc(mm), c_id.alloc(mm), c_wlo.alloc(mm), c_whi.alloc(mm),
                                              alive.alloc(mm), decided.alloc(mm)));
  size_t bytes = 0;
  long long* scal_p = scal.alloc(16);
  double* dmax = reinterpret_cast<double*>(scal_p);
  FOFB_CUDA(cub::DeviceReduce::Max(nullptr, bytes, c_R1.p, dmax, m, st));
  FOFB_CUDA(cub::DeviceReduce::Max(cub_tmp(h, bytes), bytes, c_R1.p, dmax, m, st));
  FOFB_CUDA(cub::DeviceReduce::Max(cub_tmp(h, bytes), bytes, c_th05.p, dmax + 1, m, st));
  count_launch(h, 4);
  double r2[2];
  FOFB_CUDA(cudaMemcpyAsync(r2, dmax, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
  FOFB_CUDA(cudaStreamSynchronize(st));
  R1max = r2[0];
  FOFB_REQUIRE(std::isfinite(R1max) && R1max > 0 && std::isfinite(r2[1]) && r2[1] > 0,
               "non-finite search radius (check centre redshifts)");
  big.build(h, cat, R1max);
  // packed records: the overdensity counts run on tiles (overlap.cu).  A lent list serves if its cells are large enough.
  if (smallp != &small_own && !(smallp->built && smallp->has_packed && smallp->radius >= r2[1])) smallp = &small_own;
  if (smallp == &small_own) small_own.build(h, cat, r2[1], true);

  // tracker kill targets
  {
    double* key_in = sort_keys_d.alloc(mm);
    double* key_out = sort_keys_d2.alloc(mm);
    int* val_in = sort_vals.alloc(mm);
    int* val_out = sort_vals2.alloc(mm);
    k_col4_keys<<<grid_for(m, B), B, 0, st>>>(Cn, m, key_in, val_in);
    FOFB_LAUNCH_CHECK(h);
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, key_in, key_out, val_in, val_out, m, 0, 64, st));
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, key_in, key_out, val_in, val_out, m, 0, 64, st));
    count_launch(h, 8);
    int* tcnt = t_count.alloc(mm + 2);
    FOFB_CUDA(cudaMemsetAsync(tcnt, 0, sizeof(int) * (mm + 2), st));
    FOFB_PROFILED(h, "k_kill_targets", k_kill_targets<<<grid_for(n, B), B, 0, st>>>(G, n, m, key_out, val_out, kill_target.alloc(n), tcnt, tcnt + mm));
    long long* tlen = hcap.alloc(mm + 1);
    long long* toff = t_off.alloc(mm + 1);
    k_i32_to_ll<<<grid_for(m, B), B, 0, st>>>(m, tcnt, tlen);
    FOFB_LAUNCH_CHECK(h);
    const int n_dup = d2h_scalar(h, tcnt + mm);
    // A slab run sees only its own rows: a value shared with a row of another slab would need the tracker state of
    // centres far outside the halo.  Single-process runs reproduce the by-value rule exactly.
    if (n_dup != 0 && owned)
      throw Error(FOFB_ERR_UNSUPPORTED,
                  "slab-decomposed runs need column-4 (z_upper) values that identify a galaxy row uniquely: the tracker of "
                  "fof.py:238-246 matches by value, so duplicated values couple centres across slabs; run this "
                  "catalogue on one GPU (fofb_pipeline_run / fofb_fof_run reproduce the by-value rule)");
    const long long n_trows = scan_offsets(h, tlen, toff, m);
    int* trows = t_rows.alloc((size_t)std::max<long long>(n_trows, 1));
    FOFB_CUDA(cudaMemsetAsync(tcnt, 0, sizeof(int) * mm, st));
    k_kill_rows<<<grid_for(n, B), B, 0, st>>>(n, kill_target.p, toff, tcnt, trows);
    FOFB_LAUNCH_CHECK(h);
    duplicated_keys = n_dup;
  }
  // centre cell list
  {
    cc_sdec = R1max * 1.000001;
    double cabs = 0.0;
    {  // centres may lie outside the galaxy bbox: bound |dec| by 90
      cabs = fmin(89.0, cat.dec_absmax + 1.0);
    }
    cc_sra = ra_reach(R1max, cabs);
    // bbox of centres: reuse catalogue bbox padded; clamp handles outliers
    cc_ra0 = cat.gbox.ra_min;
    cc_dec0 = cat.gbox.dec_min;
    const double ext_ra = cat.gbox.ra_max - cat.gbox.ra_min, ext_dec = cat.gbox.dec_max - cat.gbox.dec_min;
    for (;;) {
      const double cells = (floor(ext_ra / cc_sra) + 1.0) * (floor(ext_dec / cc_sdec) + 1.0);
      if (cells <= fmax(1024.0, 2.0 * (double)m) && cells < 1.0e9) break;
      cc_sra *= 1.25;
      cc_sdec *= 1.25;
    }
    cc_ncx = (int)floor(ext_ra / cc_sra) + 1;
    cc_ncy = (int)floor(ext_dec / cc_sdec) + 1;
    const int ncell = cc_ncx * cc_ncy;
    // centres by redshift, then a stable sort by cell: every cell ends up z-sorted
    double* zk_in = sort_keys_d.alloc(mm);
    double* zk_out = sort_keys_d2.alloc(mm);
    int* zi_in = sort_vals.alloc(mm);
    int* zi_out = sort_vals2.alloc(mm);
    FOFB_CUDA(cudaMemcpyAsync(zk_in, c_z.p, sizeof(double) * mm, cudaMemcpyDeviceToDevice, st));
    k_iota_i32<<<grid_for(m, B), B, 0, st>>>(m, zi_in);
    FOFB_LAUNCH_CHECK(h);
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, zk_in, zk_out, zi_in, zi_out, m, 0, 64, st));
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, zk_in, zk_out, zi_in, zi_out, m, 0, 64, st));
    count_launch(h, 8);
    unsigned* ka = cc_key_a.alloc(mm);
    unsigned* kb = cc_key_b.alloc(mm);
    int* vb = cc_idx.alloc(mm);
    k_cc_keys<<<grid_for(m, B), B, 0, st>>>(m, zi_out, c_ra.p, c_dec.p, cc_ra0, cc_dec0, 1.0 / cc_sra, 1.0 / cc_sdec, cc_ncx,
                                            cc_ncy, ka);
    FOFB_LAUNCH_CHECK(h);
    int bits = 1;
    while ((1ll << bits) < ncell) ++bits;
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, ka, kb, zi_out, vb, m, 0, bits, st));
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, ka, kb, zi_out, vb, m, 0, bits, st));
    count_launch(h, 4);
    k_cc_starts<<<grid_for(m + 1, B), B, 0, st>>>(m, kb, ncell, cc_start.alloc((size_t)ncell + 1));
    FOFB_LAUNCH_CHECK(h);
    k_gather_d<<<grid_for(m, B), B, 0, st>>>(m, vb, c_z.p, cc_z.alloc(mm));
    FOFB_LAUNCH_CHECK(h);
    k_cc_pack<<<grid_for(m, B), B, 0, st>>>(m, vb, c_ra.p, c_dec.p, c_cos.p, c_z.p, c_R1.p, c_T1.p, cc_a.alloc(mm), cc_b.alloc(mm));
    FOFB_LAUNCH_CHECK(h);
  }
  // ---- priority rounds: initial active list -------------------------------------------------------
  FOFB_CUDA(cudaMemsetAsync(resume_row.alloc(mm), 0, sizeof(int) * mm, st));
  FOFB_CUDA(cudaMemsetAsync(resume_cell.alloc(mm), 0xff, sizeof(int) * mm, st));  // -1: no scan position yet
  FOFB_CUDA(cudaMemsetAsync(resume_k.alloc(mm), 0xff, sizeof(int) * mm, st));
  int* act = active_a.alloc(mm);
  active_b.alloc(mm);
  act_in_a = true;
  // the active list walks the centre cell list, so a warp's 32 centres share cells and redshift neighbourhoods
  if (!owned) {
    FOFB_CUDA(cudaMemcpyAsync(act, cc_idx.p, sizeof(int) * mm, cudaMemcpyDeviceToDevice, st));
    na = m;
  } else {
    int* fl = flags.alloc(mm);
    k_gather_flag<<<grid_for(m, B), B, 0, st>>>(m, cc_idx.p, owned, fl);
    FOFB_LAUNCH_CHECK(h);
    na = select_flagged(h, cc_idx.p, fl, act, m, reinterpret_cast<int*>(scal_p + 8));
  }
  pool_centre.alloc(mm);
  pool_off.alloc(mm);
  pool_len.alloc(mm);
}

// Start the rounds again on the prepared catalogue with other stage-1/2/4 parameters (linking_length_factor,
// richness, overdensity): everything prepare() built depends only on max_velocity, the radii and the cosmology.
void FofState::restart_rounds(fofb_handle* h, double linking_length_factor, int richness, double overdensity) {
  cudaStream_t st = h->stream;
  if (distributed) throw Error(FOFB_ERR_STATE, "parameter sweeps are not available in distributed runs");
  params.linking_length_factor = linking_length_factor;
  params.richness = richness;
  params.overdensity = overdensity;
  ran = finished = false;
  candidates.K = merged.K = bcg.K = final_.K = 0;
  candidates.total = merged.total = bcg.total = final_.total = 0;
  pool_used = 0;
  pool_K = 0;
  rounds = 0;
  evaluated = 0;
  na = 0;
  if (m == 0) return;
  FOFB_CUDA(cudaMemsetAsync(alive.p, 1, (size_t)m, st));
  FOFB_CUDA(cudaMemsetAsync(decided.p, 0, (size_t)m, st));
  FOFB_CUDA(cudaMemsetAsync(resume_row.p, 0, sizeof(int) * (size_t)m, st));
  FOFB_CUDA(cudaMemsetAsync(resume_cell.p, 0xff, sizeof(int) * (size_t)m, st));
  FOFB_CUDA(cudaMemsetAsync(resume_k.p, 0xff, sizeof(int) * (size_t)m, st));
  FOFB_CUDA(cudaMemcpyAsync(active_a.p, cc_idx.p, sizeof(int) * (size_t)m, cudaMemcpyDeviceToDevice, st));
  act_in_a = true;
  na = m;
}

// One round: find the active centres no alive, undecided predecessor can still claim, and evaluate them.
void FofState::round_evaluate(fofb_handle* h) {
  cudaStream_t st = h->stream;
  const fofb_params& p = params;
  const int B = 256;
  if (na <= 0) return;
  const Cosmo cosmo = device_cosmo(h, p);
  Catalog& cat = cat_();
  CatalogView cv = cat.view();
  CentreCells cc{cc_ra0, cc_dec0, 1.0 / cc_sra, 1.0 / cc_sdec, cc_ncx, cc_ncy, cc_start.p, cc_idx.p, cc_z.p, cc_a.p, cc_b.p};
  CellListView clbig = big.view();
  int* act = act_in_a ? active_a.p : active_b.p;
  int* d_count = reinterpret_cast<int*>(scal.p + 8);
  {
    ++rounds;
    int* fl = flags.alloc((size_t)na);
    // FOFB_READY=lazy|ilp forces one scan kernel (diagnostics); by default the big early rounds take the lazy scan and
    // the later ones, whose few long scans are pure latency, the one with grouped loads
    static const int ready_mode = [] {
      const char* e = getenv("FOFB_READY");
      return !e ? 0 : (e[0] == 'l' ? 1 : 2);
    }();
    if (na > 8192) {
      const bool lazy = ready_mode == 1 || (ready_mode == 0 && na > (1 << 19));
      if (lazy) {
        FOFB_PROFILED(h, rounds == 1 ? "k_ready_r1" : "k_ready", k_ready_lazy<<<grid_for(na, 128), 128, 0, st>>>(na, act, cc, c_ra.p, c_dec.p, c_cos.p, c_z.p, c_R1.p, c_T1.p, alive.p, decided.p,
                                               t_off.p, t_rows.p, cat.ra_row.p, cat.dec_row.p, cos_row.p, cat.z_row.p,
                                               p.max_velocity, R1max, fl, resume_row.p, resume_cell.p, resume_k.p));
      } else {
        FOFB_PROFILED(h, rounds == 1 ? "k_ready_r1" : "k_ready", k_ready<<<grid_for(na, kReadyThreads), kReadyThreads, 0, st>>>(na, act, cc, c_ra.p, c_dec.p, c_cos.p, c_z.p, c_R1.p, c_T1.p, alive.p, decided.p,
                                               t_off.p, t_rows.p, cat.ra_row.p, cat.dec_row.p, cos_row.p, cat.z_row.p,
                                               p.max_velocity, R1max, fl, resume_row.p, resume_cell.p, resume_k.p));
      }
    } else {
      FOFB_PROFILED(h, "k_ready_warp", k_ready_warp<<<grid_for((int64_t)na * 32, 128), 128, 0, st>>>(na, act, cc, c_ra.p, c_dec.p, c_cos.p, c_z.p, c_R1.p, c_T1.p, alive.p, decided.p,
                                               t_off.p, t_rows.p, cat.ra_row.p, cat.dec_row.p, cos_row.p, cat.z_row.p,
                                               p.max_velocity, R1max, fl));
    }
    int* rdy = ready.alloc((size_t)na);
    const int nr_all = select_flagged(h, act, fl, rdy, na, d_count);
    if (nr_all == 0 && !distributed) throw Error(FOFB_ERR_INTERNAL, "priority rounds made no progress");
    // evaluate the ready centres in bounded batches
    const int kBatch = 1 << 16;
    for (int b0 = 0; b0 < nr_all; b0 += kBatch) {
      const int nr = std::min(kBatch, nr_all - b0);
      const int* rb = rdy + b0;
      evaluated += nr;
      const int wblocks = grid_for((int64_t)nr * 32, 128);
      int* nv_p = nv.alloc(nr);
      double4* bb1 = bbox1.alloc(nr);
      FOFB_PROFILED(h, "k_virial_count", k_virial<0><<<wblocks, 128, 0, st>>>(nr, rb, cv, clbig, c_ra.p, c_dec.p, c_cos.p, c_R1.p, c_wlo.p, c_whi.p,
                                           p.grid_cells, p.min_virial, nv_p, bb1, nullptr, nullptr, nullptr));
      long long* need = hcap.alloc((size_t)nr + 1);
      long long* vo = voff.alloc((size_t)nr + 1);
      k_alloc_sizes<<<grid_for(nr, B), B, 0, st>>>(nr, nv_p, p.min_virial, need);
      FOFB_LAUNCH_CHECK(h);
      const long long total_v = scan_offsets(h, need, vo, nr);
      FOFB_REQUIRE(total_v < (1ll << 31), "virial scratch exceeds 2^31 entries in one batch");
      int* nF_p = nF.alloc(nr);
      int* nM_p = nM.alloc(nr);
      int* pass_p = pass.alloc(nr);
      long long* ho = hoff.alloc((size_t)nr + 1);
      const size_t tv = (size_t)std::max<long long>(total_v, 1);
      int* v2 = v2row.alloc(tv);
      int* mr = mrows.alloc(tv);
      int* nml = nm_list.alloc(tv);
      int* fbatch = first_batch.alloc(tv);
      double* LLp = LL.alloc(nr);
      double4* bb2 = bbox2.alloc(nr);
      unsigned long long* k1 = keys1.alloc(tv);
      unsigned long long* k1s = keys1b.alloc(tv);
      unsigned long long* k2 = keys2.alloc(tv);
      unsigned long long* k2s = keys2b.alloc(tv);
      long long total_h = 0;
      // (friend flag, friends' cell, virial cell, row) fits one 64-bit key up to 181 grid cells per axis: then the
      // virial points need no sort of their own
      int key_bits = 1;
      while ((1ll << key_bits) < (long long)p.grid_cells * p.grid_cells) ++key_bits;
      if (1 + 2 * key_bits + 32 > 64) key_bits = 0;
      const bool radix = key_bits > 0 && total_v > radix_sort_above();
      int* seg_a = radix ? seg_of_a.alloc(tv) : nullptr;
      int* seg_b = radix ? seg_of_b.alloc(tv) : nullptr;
      const unsigned long long* k2sorted = radix ? k2 : k2s;
      if (total_v > 0) {
        int* cur = cursor.alloc(nr);
        FOFB_CUDA(cudaMemsetAsync(cur, 0, sizeof(int) * nr, st));
        FOFB_PROFILED(h, "k_virial_fill", k_virial<1><<<wblocks, 128, 0, st>>>(nr, rb, cv, clbig, c_ra.p, c_dec.p, c_cos.p, c_R1.p, c_wlo.p, c_whi.p,
                                             p.grid_cells, p.min_virial, nv_p, bb1, vo, cur, k1));
        if (!key_bits) segmented_sort(h, k1, k1s, total_v, nr, vo);
        const unsigned long long* k1v = key_bits ? k1 : k1s;
        FOFB_PROFILED(h, "k_hop1", k_hop1<<<wblocks, 128, 0, st>>>(nr, rb, nv_p, vo, k1v, cat.ra_row.p, cat.dec_row.p, cos_row.p, c_ra.p, c_dec.p,
                                        c_cos.p, c_z.p, cosmo, p, LLp, bb2, nF_p, k2, key_bits, radix ? seg_a : nullptr));
        if (radix) {
          // two stable device-wide radix sorts (by key, then by centre) instead of one segmented sort: at millions of
          // keys in segments of a few hundred the onesweep passes are several times faster
          size_t bytes = 0;
          FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, k2, k2s, seg_a, seg_b, total_v, 0, 64, st));
          FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, k2, k2s, seg_a, seg_b, total_v, 0, 64, st));
          int seg_bits = 1;
          while ((1 << seg_bits) < nr) ++seg_bits;
          FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, seg_b, seg_a, k2s, k2, total_v, 0, seg_bits, st));
          FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, seg_b, seg_a, k2s, k2, total_v, 0, seg_bits, st));
          count_launch(h, 12);
        } else {
          segmented_sort(h, k2, k2s, total_v, nr, vo);
        }
        k_hash_sizes<<<grid_for(nr, B), B, 0, st>>>(nr, nv_p, nF_p, p.min_virial, need);
        FOFB_LAUNCH_CHECK(h);
        total_h = scan_offsets(h, need, ho, nr);
        if (total_h > 0) {
          unsigned long long* tab = htab.alloc((size_t)total_h);
          FOFB_CUDA(cudaMemsetAsync(tab, 0xff, sizeof(unsigned long long) * (size_t)total_h, st));
        }
      } else {
        FOFB_CUDA(cudaMemsetAsync(nF_p, 0, sizeof(int) * nr, st));
        FOFB_CUDA(cudaMemsetAsync(ho, 0, sizeof(long long) * ((size_t)nr + 1), st));
      }
      if (total_v > 0) {
        int* own = owner.alloc(tv);
        double* v2a = v2ra.alloc(tv);
        double* v2d = v2dec.alloc(tv);
        double* v2c = v2cos.alloc(tv);
        unsigned char* nov = ismember.alloc(tv);
        k_hop2_setup<<<wblocks, 128, 0, st>>>(nr, nv_p, nF_p, vo, k1s, k2sorted, cat.ra_row.p, cat.dec_row.p, cos_row.p, p.min_virial,
                                              v2, own, v2a, v2d, v2c, mr, key_bits);
        FOFB_LAUNCH_CHECK(h);
        if (total_h > 0) {
          FOFB_PROFILED(h, "k_hop2_insert", k_hop2_insert<<<grid_for(total_v, B), B, 0, st>>>(total_v, own, vo, nF_p, ho, need, htab.p, G, v2));
          FOFB_PROFILED(h, "k_hop2_novel", k_hop2_novel<<<grid_for(total_v, B), B, 0, st>>>(total_v, own, vo, nF_p, ho, need, htab.p, G, v2, nov));
        }
      }
      {
        int* qc = hop2_qcount.alloc(4);
        int* ql = hop2_qlist.alloc((size_t)nr * 3);
        int* nnm = hop2_nnm.alloc(nr);
        int* nms = nm_sorted.alloc(tv);
        FOFB_CUDA(cudaMemsetAsync(qc, 0, sizeof(int) * 4, st));
        const int cap_rows = hop2_smem_rows();
        const int small_rows = std::min(kHop2SmallRows, cap_rows), large_rows = std::min(kHop2LargeRows, cap_rows);
        FOFB_PROFILED(h, "k_hop2_front", k_hop2_front<<<wblocks, 128, 0, st>>>(nr, rb, nv_p, nF_p, vo, need, ismember.p, p, nml, nnm, mr, nM_p,
                                                    pass_p, kill_target.p, alive.p, decided.p, small_rows, large_rows, qc, ql));
        static bool attr_set[64] = {};
        if (!attr_set[h->device & 63]) {
          FOFB_CUDA(cudaFuncSetAttribute(k_hop2_heavy<kHop2SmallRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hop2_smem_bytes(kHop2SmallRows)));
          FOFB_CUDA(cudaFuncSetAttribute(k_hop2_heavy<kHop2LargeRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hop2_smem_bytes(kHop2LargeRows)));
          attr_set[h->device & 63] = true;
        }
        // persistent CTAs over the three queues: <= 256 candidates (4 CTAs per SM), <= 1024 (one per SM), the rest
        // (global-memory path).  The queues are independent, and a round waits for its slowest centre: the second and
        // third queue run on side streams beside the first.  An empty queue costs one launch.
        const int g0 = std::min(nr, h->sm_count * 4), g1 = std::min(nr, h->sm_count), g2 = std::min(nr, h->sm_count * 2);
        if (!hop2_side[0]) {
          for (int q = 0; q < 2; ++q) {
            FOFB_CUDA(cudaStreamCreateWithFlags(&hop2_side[q], cudaStreamNonBlocking));
            FOFB_CUDA(cudaEventCreateWithFlags(&hop2_done[q], cudaEventDisableTiming));
          }
          FOFB_CUDA(cudaEventCreateWithFlags(&hop2_go, cudaEventDisableTiming));
        }
        FOFB_CUDA(cudaEventRecord(hop2_go, st));
        FOFB_CUDA(cudaStreamWaitEvent(hop2_side[0], hop2_go, 0));
        FOFB_CUDA(cudaStreamWaitEvent(hop2_side[1], hop2_go, 0));
        k_hop2_heavy<kHop2LargeRows><<<g1, kHop2Threads, hop2_smem_bytes(kHop2LargeRows), hop2_side[0]>>>(
            1, nr, qc, ql, rb, nF_p, nnm, vo, ho, need, htab.p, G, v2ra.p, v2dec.p, v2cos.p, LLp, bb2, p, v2, mr, nml, fbatch, nms, nM_p,
            pass_p, kill_target.p, alive.p, decided.p, k1, hop2_sort_above());
        FOFB_LAUNCH_CHECK(h);
        FOFB_CUDA(cudaEventRecord(hop2_done[0], hop2_side[0]));
        k_hop2_heavy<0><<<g2, kHop2Threads, 0, hop2_side[1]>>>(
            2, nr, qc, ql, rb, nF_p, nnm, vo, ho, need, htab.p, G, v2ra.p, v2dec.p, v2cos.p, LLp, bb2, p, v2, mr, nml, fbatch, nms, nM_p,
            pass_p, kill_target.p, alive.p, decided.p, k1, hop2_sort_above());
        FOFB_LAUNCH_CHECK(h);
        FOFB_CUDA(cudaEventRecord(hop2_done[1], hop2_side[1]));
        FOFB_PROFILED(h, "k_hop2", k_hop2_heavy<kHop2SmallRows><<<g0, kHop2Threads, hop2_smem_bytes(kHop2SmallRows), st>>>(
            0, nr, qc, ql, rb, nF_p, nnm, vo, ho, need, htab.p, G, v2ra.p, v2dec.p, v2cos.p, LLp, bb2, p, v2, mr, nml, fbatch, nms, nM_p,
            pass_p, kill_target.p, alive.p, decided.p, k1, hop2_sort_above()));
        FOFB_CUDA(cudaStreamWaitEvent(st, hop2_done[0], 0));
        FOFB_CUDA(cudaStreamWaitEvent(st, hop2_done[1], 0));
      }
      // commit the clusters that passed the richness gate
      long long* len = need;  // reuse
      long long* dst = ho;    // reuse
      k_commit_sizes<<<grid_for(nr, B), B, 0, st>>>(nr, pass_p, nM_p, len);
      FOFB_LAUNCH_CHECK(h);
      const long long add = scan_offsets(h, len, dst, nr);
      if (add > 0) {
        long long* pl = moff.alloc((size_t)nr + 1);
        long long* slot = reinterpret_cast<long long*>(keys2.alloc(tv > (size_t)nr + 1 ? tv : (size_t)nr + 1));
        k_pass_to_ll<<<grid_for(nr, B), B, 0, st>>>(nr, pass_p, pl);
        FOFB_LAUNCH_CHECK(h);
        const long long newK = scan_offsets(h, pl, slot, nr);
        // grow the pool, preserving its contents
        if ((size_t)(pool_used + add) > pool_rows.cap) {
          DBuf<int> bigger;
          bigger.alloc((size_t)(pool_used + add) * 2 + 1024);
          if (pool_used) FOFB_CUDA(cudaMemcpyAsync(bigger.p, pool_rows.p, sizeof(int) * pool_used, cudaMemcpyDeviceToDevice, st));
          FOFB_CUDA(cudaStreamSynchronize(st));
          std::swap(bigger.p, pool_rows.p);
          std::swap(bigger.cap, pool_rows.cap);
        }
        k_commit<<<wblocks, 128, 0, st>>>(nr, rb, pass_p, nM_p, vo, mr, dst, slot, pool_used, pool_K, pool_rows.p,
                                          pool_centre.p, pool_off.p, pool_len.p);
        FOFB_LAUNCH_CHECK(h);
        pool_used += add;
        pool_K += (int)newK;
      }
    }
  }
}

// After the round's kills (and, when distributed, after the state arrays were merged across processes).
void FofState::round_advance(fofb_handle* h) {
  cudaStream_t st = h->stream;
  if (na <= 0) return;
  int* act = act_in_a ? active_a.p : active_b.p;
  int* act_next = act_in_a ? active_b.p : active_a.p;
  int* fl = flags.alloc((size_t)na);
  k_still_active<<<grid_for(na, 256), 256, 0, st>>>(na, act, alive.p, decided.p, fl);
  FOFB_LAUNCH_CHECK(h);
  na = select_flagged(h, act, fl, act_next, na, reinterpret_cast<int*>(scal.p + 8));
  act_in_a = !act_in_a;
}

// ---- tracker state for the one-collective-per-round exchange (0 alive and undecided, 1 switched off, 2 evaluated) ----
__global__ void k_state_pack(long long n, const long long* lpos, const long long* bpos, const unsigned char* alive,
                             const unsigned char* decided, unsigned char* buf) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long i = lpos[t];
  buf[bpos[t]] = decided[i] ? 2 : (alive[i] ? 0 : 1);
}

__global__ void k_state_unpack(long long n, const long long* lpos, const long long* bpos, const unsigned char* buf,
                               unsigned char* alive, unsigned char* decided) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const long long i = lpos[t];
  const unsigned char s = buf[bpos[t]];
  if (s == 2) decided[i] = 1;
  else if (s == 1) alive[i] = 0;
}

void dist_state_pack(fofb_handle* h, FofState& S, long long n_shared, const long long* lpos, const long long* bpos, unsigned char* buf,
                     long long n_buf) {
  cudaStream_t st = h->stream;
  FOFB_CUDA(cudaMemsetAsync(buf, 0, (size_t)n_buf, st));
  if (S.na > 0) FOFB_CUDA(cudaMemsetAsync(buf + (n_buf - 1), 1, 1, st));
  if (n_shared > 0) {
    k_state_pack<<<grid_for(n_shared, 256), 256, 0, st>>>(n_shared, lpos, bpos, S.alive.p, S.decided.p, buf);
    FOFB_LAUNCH_CHECK(h);
  }
}

void dist_state_unpack(fofb_handle* h, FofState& S, long long n_shared, const long long* lpos, const long long* bpos,
                       const unsigned char* buf) {
  if (n_shared <= 0) return;
  k_state_unpack<<<grid_for(n_shared, 256), 256, 0, h->stream>>>(n_shared, lpos, bpos, buf, S.alive.p, S.decided.p);
  FOFB_LAUNCH_CHECK(h);
}

void FofState::finalize(fofb_handle* h) {
  cudaStream_t st = h->stream;
  const int B = 256;
  const size_t mm = (size_t)std::max(m, 1);
  size_t bytes = 0;
  // ---- candidates in priority order (= the order the sequential loop appends them, fof.py:232) ------
  candidates.K = pool_K;
  candidates.total = pool_used;
  if (pool_K > 0) {
    const size_t K = (size_t)pool_K;
    int* slot_in = sort_vals.alloc(std::max(K, mm));
    int* slot_out = sort_vals2.alloc(std::max(K, mm));
    int* key_out = flags.alloc(std::max(K, mm));
    k_iota_i32<<<grid_for(pool_K, B), B, 0, st>>>(pool_K, slot_in);
    FOFB_LAUNCH_CHECK(h);
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, pool_centre.p, key_out, slot_in, slot_out, pool_K, 0, 32, st));
    FOFB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp(h, bytes), bytes, pool_centre.p, key_out, slot_in, slot_out, pool_K, 0, 32, st));
    count_launch(h, 6);
    long long* len = hcap.alloc(K + 1);
    long long* off = candidates.off.alloc(K + 1);
    k_list_sizes<<<grid_for(pool_K, B), B, 0, st>>>(pool_K, slot_out, pool_len.p, len);
    FOFB_LAUNCH_CHECK(h);
    const long long tot = scan_offsets(h, len, off, pool_K);
    if (tot != pool_used) throw Error(FOFB_ERR_INTERNAL, "candidate pool accounting mismatch");
    k_list_gather<<<grid_for((int64_t)pool_K * 32, 128), 128, 0, st>>>(pool_K, slot_out, pool_centre.p, pool_off.p, pool_len.p,
                                                                       pool_rows.p, off, candidates.centre.alloc(K),
                                                                       candidates.rows.alloc((size_t)std::max<long long>(tot, 1)));
    FOFB_LAUNCH_CHECK(h);
  }
}

}  // namespace fofb
